mkdir -p gpurun_out/r14
for lib in "" ns3 ns4; do for mib in 8 16 32 64; do
  L=""; [ -n "$lib" ] && L=$PWD/glfer_b200/libglfer_b200_$lib.so
  GLFER_B200_LIB=$L GLFER_B200_CHUNK_MIB=$mib python bench.py --steps 20 --warmup 3 --no-cpu > gpurun_out/r14/e2e_${lib:-ns2}_$mib.json 2>/dev/null
  echo "slots=${lib:-ns2} chunk=$mib MiB: $(python -c "import json;d=json.load(open('gpurun_out/r14/e2e_${lib:-ns2}_$mib.json'));print('%.4g frames/s %.2f ms'%(d['e2e']['value'],d['e2e']['ms_per_step']))")"
done; done
