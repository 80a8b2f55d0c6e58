"""Host-side limit of concurrent H2D + D2H copies on several GPUs of one box: aggregate GB/s for pinned memory
from cudaHostAlloc (default / write-combined input) and for transparent-huge-page memory pinned with
cudaHostRegister.  One host thread drives all devices (copies are asynchronous)."""
import ctypes as C
import mmap
import sys
import time

import torch

rt = C.CDLL("libcudart.so.12") if True else None
rt.cudaHostAlloc.argtypes = [C.POINTER(C.c_void_p), C.c_size_t, C.c_uint]
rt.cudaHostRegister.argtypes = [C.c_void_p, C.c_size_t, C.c_uint]
rt.cudaMemcpyAsync.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_int, C.c_void_p]
libc = C.CDLL("libc.so.6", use_errno=True)
libc.mmap.restype = C.c_void_p
libc.mmap.argtypes = [C.c_void_p, C.c_size_t, C.c_int, C.c_int, C.c_int, C.c_long]
libc.madvise.argtypes = [C.c_void_p, C.c_size_t, C.c_int]
NB = 512 << 20


def host_default():
    p = C.c_void_p()
    assert rt.cudaHostAlloc(C.byref(p), NB, 0) == 0
    return p.value


def host_wc():
    p = C.c_void_p()
    assert rt.cudaHostAlloc(C.byref(p), NB, 4) == 0      # cudaHostAllocWriteCombined
    return p.value


def host_thp():
    size = NB + (2 << 20)
    base = libc.mmap(None, size, 3, 0x22, -1, 0)           # PROT_READ|WRITE, MAP_PRIVATE|MAP_ANONYMOUS
    assert base not in (None, C.c_void_p(-1).value)
    al = (base + (2 << 20) - 1) & ~((2 << 20) - 1)
    rc = libc.madvise(al, NB, 14)                          # MADV_HUGEPAGE
    C.memset(al, 1, NB)                                    # fault the pages in (as huge pages when THP allows)
    assert rt.cudaHostRegister(al, NB, 0) == 0, "cudaHostRegister"
    return al, rc


def run(ndev, kind):
    src, dst, dev_in, dev_out, streams = [], [], [], [], []
    for d in range(ndev):
        torch.cuda.set_device(d)
        if kind == "default":
            a, b = host_default(), host_default()
        elif kind == "wc_in":
            a, b = host_wc(), host_default()
        else:
            (a, _), (b, _) = host_thp(), host_thp()
        C.memset(a, 1, NB)
        src.append(a)
        dst.append(b)
        dev_in.append(torch.empty(NB, dtype=torch.uint8, device=f"cuda:{d}"))
        dev_out.append(torch.ones(NB, dtype=torch.uint8, device=f"cuda:{d}"))
        streams.append((torch.cuda.Stream(d), torch.cuda.Stream(d)))

    def once():
        for d in range(ndev):
            torch.cuda.set_device(d)
            rt.cudaMemcpyAsync(dev_in[d].data_ptr(), src[d], NB, 1, streams[d][0].cuda_stream)
            rt.cudaMemcpyAsync(dst[d], dev_out[d].data_ptr(), NB, 2, streams[d][1].cuda_stream)

    def sync():
        for d in range(ndev):
            torch.cuda.synchronize(d)

    once(); sync()
    t0 = time.perf_counter()
    for _ in range(5):
        once()
    sync()
    dt = (time.perf_counter() - t0) / 5
    print(f"{ndev} GPUs, {kind:8s}: {NB / dt / 1e9:6.1f} GB/s each way per GPU, {2 * ndev * NB / dt / 1e9:6.1f} GB/s in total", flush=True)


if __name__ == "__main__":
    n = torch.cuda.device_count()
    try:
        print("THP:", open("/sys/kernel/mm/transparent_hugepage/enabled").read().strip())
    except OSError as e:
        print("THP: ?", e)
    for ndev in sorted({1, min(2, n), n}):
        for kind in ("default", "wc_in", "thp"):
            try:
                run(ndev, kind)
            except AssertionError as e:
                print(ndev, kind, "failed:", e)
