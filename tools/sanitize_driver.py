"""Small invocations of every kernel family for compute-sanitizer (memcheck / racecheck / synccheck):

    compute-sanitizer --tool memcheck python tools/sanitize_driver.py

Sizes are chosen so that every frame group of the resident grid walks several frames (the TMA ring
slots are then re-filled and re-read: the hand-placed fence.proxy.async + barrier (A) hand-over is
what racecheck is asked to look at) while the whole run stays short under the tool's slow-down."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from glfer_b200 import api, synth      # noqa: E402


def main():
    quick = "--quick" in sys.argv
    x = synth.qrss_stream(2048 * (1400 if quick else 3000), fs=48000, seed=11, dot_s=0.2)
    cases = [
        ("ring 50% (QSC=3)", dict(n=4096, window_type=0, overlap=0.5, sub_mean=True), 0),
        ("ring 75% (QSC=2) + plain averaging", dict(n=4096, window_type=7, overlap=0.75, sub_mean=True, avg_mode=2, avg_depth=4,
                                                    avg_minbin=34, avg_maxbin=102), 0),
        ("ring 87.5% (run-time geometry)", dict(n=2048, window_type=1, overlap=0.875, sub_mean=True), 0),
        ("ring, small frames (8 groups per CTA)", dict(n=512, window_type=0, overlap=0.5, sub_mean=True), 0),
        ("ring multitaper", dict(n=4096, mode=1, overlap=0.5, sub_mean=True, mtm_w=4.0, mtm_kmax=3), 0),
        ("ring N=16384", dict(n=16384, window_type=0, overlap=0.5, sub_mean=True), 0),
        ("ring N=32768 (table twiddles)", dict(n=32768, window_type=0, overlap=0.5, sub_mean=True), 0),
        ("general multitaper N=32768", dict(n=32768, mode=1, overlap=0.5, sub_mean=True, mtm_w=8.0, mtm_kmax=3), 0),
        ("general, odd hop + RA9MB + limiter", dict(n=4096, window_type=0, overlap=0.9, sub_mean=True, a=0.01, limiter=1), 0),
        ("general, zeroed history", dict(n=1024, window_type=0, overlap=0.75, sub_mean=False, zero_history=True), 0),
        ("pair kernel", dict(n=4096, window_type=0, overlap=0.5, sub_mean=True), 4),
        ("warp-per-frame kernel", dict(n=4096, window_type=0, overlap=0.5, sub_mean=True), 3),
        ("sumavg averaging", dict(n=1024, window_type=0, overlap=0.5, sub_mean=True, avg_mode=1, avg_depth=3, avg_minbin=10,
                                  avg_maxbin=200, avg_max0=1), 0),
        ("deep averaging (sliding kernel)", dict(n=1024, window_type=0, overlap=0.5, sub_mean=True, avg_mode=3, avg_depth=40,
                                               avg_minbin=10, avg_maxbin=200), 0),
        ("LMP", dict(n=2048, mode=3, overlap=0.5, sub_mean=True, lmp_av=4), 0),
    ]
    for name, kw, pref in cases:
        api.set_kernel_preference(pref)
        p = api.GramPlan(**kw)
        r = p.run(x)
        ok = np.isfinite(r["psd"]).all()
        print(f"{name}: frames {r['psd'].shape[0]} family {api.last_kernel_family()} finite {ok}", flush=True)
        assert ok
        p.close()
    api.set_kernel_preference(0)
    p = api.GramPlan(n=2048, window_type=0, overlap=0.5, sub_mean=True)
    d = p.run_display(x[: 2048 * 300], log_scale=True, autoscale=True, want_rgb=True,
                      colortab=np.repeat(np.arange(256, dtype=np.uint8), 3))
    print("display", d["levels"].shape, int(d["levels"].sum()), flush=True)
    print("launches", api.kernel_launches())


if __name__ == "__main__":
    main()
