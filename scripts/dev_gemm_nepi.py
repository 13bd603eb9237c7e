import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.chdir(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
src = open("scripts/dev_gemm_epi.py").read().split("cases = [")[0]
exec(src)
for nepi in (8, 16):
    lib.mdv_gemm_tune(0, nepi << 8, 0)
    for M, N, K, kind in [(524288, 512, 64, "fc2d_new"), (524288, 512, 64, "fc1_new"), (524288, 64, 512, "res"), (131072, 1024, 128, "fc2d_new"), (131072, 1024, 128, "fc1_new")]:
        fn, keep = make(M, N, K, kind)
        print(f"nepi={nepi} M={M} N={N} K={K} {kind}: {bench(fn, 10):.1f} us", flush=True)
        del keep
