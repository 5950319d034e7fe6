import sys, os, subprocess
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if len(sys.argv) > 1:
    import torch
    from jorldy_b200._lib import C
    torch.manual_seed(0)
    for (M, N, K) in [(128, 128, 32), (128, 128, 512), (4096, 512, 512)]:
        x = torch.randn(M, K, device="cuda"); w = torch.randn(N, K, device="cuda") / K ** 0.5; b = torch.randn(N, device="cuda")
        y = torch.full((M, N), float("nan"), device="cuda")
        C.jb_linear_fwd_tc(x.data_ptr(), w.data_ptr(), b.data_ptr(), y.data_ptr(), M, K, N, 0, torch.cuda.current_stream().cuda_stream)
        torch.cuda.synchronize()
        ref = (x.double() @ w.double().t() + b.double()).float()
        err = (y - ref).abs().max().item(); rel = err / ref.abs().max().item()
        print(f"variant {os.environ.get('JB_TC_VARIANT','0')} {M}x{N}x{K}: max abs err {err:.3e} rel {rel:.3e} nan={torch.isnan(y).sum().item()}", flush=True)
    # timing
    M, N, K = 4096, 512, 512
    x = torch.randn(M, K, device="cuda"); w = torch.randn(N, K, device="cuda"); b = torch.randn(N, device="cuda"); y = torch.empty(M, N, device="cuda")
    cs = lambda: torch.cuda.current_stream().cuda_stream
    for fn, name in [(lambda: C.jb_linear_fwd_tc(x.data_ptr(), w.data_ptr(), b.data_ptr(), y.data_ptr(), M, K, N, 1, cs()), "tcgen05 3xTF32"),
                     (lambda: C.jb_linear_fwd(x.data_ptr(), w.data_ptr(), b.data_ptr(), y.data_ptr(), M, K, N, 1, cs()), "fp32 FFMA")]:
        g = torch.cuda.CUDAGraph()
        fn(); torch.cuda.synchronize()
        with torch.cuda.graph(g):
            for _ in range(20): fn()
        g.replay(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
        t = e0.elapsed_time(e1) / 20
        print(f"{name}: {t*1000:.1f} us  ({2*M*N*K/t/1e9:.1f} TFLOP/s)", flush=True)
else:
    for v in ["0"]:
        r = subprocess.run([sys.executable, __file__, "child"], env=dict(os.environ, JB_TC_VARIANT=v), capture_output=True, text=True, timeout=150)
        print(r.stdout[-1500:], r.stderr[-600:] if r.returncode else "")
