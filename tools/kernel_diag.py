"""First-contact diagnostics for the CUDA kernels on a real B200 (run under gpurun).

Each case runs in its own subprocess under a timeout so a dead-locked kernel cannot hang the box.
Prints max-abs / relative error against a torch fp32 computation of the same op on the GPU.
"""
import os
os.environ.setdefault("SABER_B200_ALLOW_RANDOM_INIT", "1")  # probes run on synthetic random-init weights
import subprocess, sys, os, json

CASES = [
    ("gemm", dict(M=256, N=128, K=64, bn=128)),
    ("gemm", dict(M=256, N=128, K=128, bn=128)),
    ("gemm", dict(M=128, N=64, K=64, bn=64)),
    ("gemm", dict(M=128, N=256, K=64, bn=256)),
    ("gemm", dict(M=4096, N=1728, K=576, bn=0)),
    ("gemm", dict(M=4096, N=2304, K=576, bn=0, act=1, bias=1)),
    ("gemm", dict(M=4096, N=576, K=2304, bn=0, bias=1, res=1)),
    ("gemm", dict(M=65536, N=432, K=144, bn=0, bias=1)),
    ("gemm", dict(M=65536, N=144, K=160, bn=0, bias=1, res=1, res_mod=65536 // 2)),
    ("gemm", dict(M=1000, N=100, K=72, bn=0, bias=1, out_f32=1)),
    ("gemm", dict(M=512, N=4, K=256, bn=0, bias=1, out_f32=1)),
    ("ln", dict(M=4096, C=576)),
    ("ln", dict(M=1000, C=144, in_f32=1)),
    ("attn", dict(B=2, heads=8, hd=72, nq=4096, nk=4096)),
    ("attn", dict(B=3, heads=8, hd=16, nq=7, nk=4096)),
    ("attn", dict(B=3, heads=8, hd=16, nq=4096, nk=7)),
    ("attn", dict(B=3, heads=8, hd=32, nq=8, nk=8)),
    ("wattn", dict(B=2, H=64, W=64, heads=8, hd=72, ws=16, pool=1)),
    ("wattn", dict(B=2, H=256, W=256, heads=2, hd=72, ws=8, pool=1)),
    ("wattn", dict(B=1, H=256, W=256, heads=4, hd=72, ws=8, pool=2)),
    ("wattn", dict(B=1, H=64, W=64, heads=4, hd=96, ws=14, pool=1)),
    ("wattn", dict(B=1, H=64, W=64, heads=8, hd=96, ws=14, pool=2)),
    ("wattn", dict(B=1, H=64, W=64, heads=8, hd=72, ws=0, pool=1)),
    ("misc", dict()),
]


def run_case(kind, kw):
    import torch
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from saber_b200 import ops
    torch.manual_seed(0)
    dev = "cuda"
    if kind == "gemm":
        M, N, K = kw["M"], kw["N"], kw["K"]
        a = torch.randn(M, K, device=dev).to(torch.bfloat16)
        w = (torch.randn(N, K, device=dev) / K ** 0.5).to(torch.bfloat16)
        bias = torch.randn(N, device=dev) if kw.get("bias") else None
        res = None
        res_mod = kw.get("res_mod", 0)
        if kw.get("res"):
            res = torch.randn(res_mod if res_mod else M, N, device=dev)
        out_dtype = torch.float32 if kw.get("out_f32") else torch.bfloat16
        out = ops.gemm(a, w, bias, kw.get("act", 0), res, res_mod, out_dtype, force_bn=kw["bn"])
        torch.cuda.synchronize()
        ref = a.float() @ w.float().t()
        if bias is not None:
            ref = ref + bias
        if kw.get("act") == 1:
            ref = torch.nn.functional.gelu(ref)
        if res is not None:
            ref = ref + (res.repeat(M // res.shape[0], 1) if res_mod else res)
        err = (out.float() - ref).abs()
        rel = err.max().item() / ref.abs().max().item()
        bad = (err > 0.05 * ref.abs().max()).nonzero()
        msg = f"max_abs={err.max().item():.4g} rel={rel:.3g} nbad={bad.shape[0]}"
        if bad.shape[0]:
            msg += f" first_bad={bad[:4].tolist()} rows_bad={bad[:,0].unique()[:8].tolist()} cols_bad={bad[:,1].unique()[:8].tolist()}"
        ok = rel < 2e-2
        # timing
        for _ in range(3):
            ops.gemm(a, w, bias, kw.get("act", 0), res, res_mod, out_dtype, force_bn=kw["bn"])
        e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
        e0.record()
        for _ in range(20):
            ops.gemm(a, w, bias, kw.get("act", 0), res, res_mod, out_dtype, force_bn=kw["bn"])
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 20
        msg += f" time={ms*1e3:.1f}us tflops={2*M*N*K/ms/1e9:.1f}"
        return ok, msg
    if kind == "ln":
        M, Cc = kw["M"], kw["C"]
        x = torch.randn(M, Cc, device=dev) * 3 + 1
        if not kw.get("in_f32"):
            x = x.to(torch.bfloat16)
        g = torch.randn(Cc, device=dev); b = torch.randn(Cc, device=dev)
        out = ops.layernorm(x, g, b, 1e-6, torch.float32)
        ref = torch.nn.functional.layer_norm(x.float(), (Cc,), g, b, 1e-6)
        err = (out - ref).abs().max().item()
        return err < 1e-4, f"max_abs={err:.3g}"
    if kind == "attn":
        B, h, hd, nq, nk = kw["B"], kw["heads"], kw["hd"], kw["nq"], kw["nk"]
        q = torch.randn(B * nq, h * hd, device=dev).to(torch.bfloat16)
        k = torch.randn(B * nk, h * hd, device=dev).to(torch.bfloat16)
        v = torch.randn(B * nk, h * hd, device=dev).to(torch.bfloat16)
        out = ops.attention(q, k, v, B, h, nq, nk)
        torch.cuda.synchronize()
        qf = q.float().view(B, nq, h, hd).transpose(1, 2)
        kf = k.float().view(B, nk, h, hd).transpose(1, 2)
        vf = v.float().view(B, nk, h, hd).transpose(1, 2)
        ref = torch.nn.functional.scaled_dot_product_attention(qf, kf, vf).transpose(1, 2).reshape(B * nq, h * hd)
        err = (out.float() - ref).abs().max().item()
        e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
        for _ in range(3): ops.attention(q, k, v, B, h, nq, nk)
        e0.record()
        for _ in range(10): ops.attention(q, k, v, B, h, nq, nk)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
        return err < 3e-2, f"max_abs={err:.3g} time={ms*1e3:.1f}us tflops={4*B*h*nq*nk*hd/ms/1e9:.1f}"
    if kind == "wattn":
        import torch.nn.functional as F
        B, H, W, h, hd, ws, pool = kw["B"], kw["H"], kw["W"], kw["heads"], kw["hd"], kw["ws"], kw["pool"]
        Cc = h * hd
        bias = torch.randn(3 * Cc, device=dev)
        qkv = (torch.randn(B * H * W, 3 * Cc, device=dev) + bias).to(torch.bfloat16)
        out = ops.window_attention(qkv, bias, B, H, W, h, ws, pool)
        torch.cuda.synchronize()
        # reference: upstream window_partition semantics on the qkv tensor
        x = qkv.float().view(B, H, W, 3 * Cc)
        wsz = ws if ws > 0 else H
        ph, pw = (wsz - H % wsz) % wsz, (wsz - W % wsz) % wsz
        if ph or pw:
            padv = bias.to(torch.bfloat16).float().view(1, 1, 1, -1)
            xp = padv.expand(B, H + ph, W + pw, 3 * Cc).clone()
            xp[:, :H, :W] = x
            x = xp
        Hp, Wp = H + ph, W + pw
        xw = x.view(B, Hp // wsz, wsz, Wp // wsz, wsz, 3 * Cc).permute(0, 1, 3, 2, 4, 5).reshape(-1, wsz, wsz, 3 * Cc)
        nW = xw.shape[0]
        qkv_w = xw.reshape(nW, wsz * wsz, 3, h, hd)
        q, k, v = qkv_w.unbind(2)
        if pool > 1:
            q = q.reshape(nW, wsz, wsz, Cc).permute(0, 3, 1, 2)
            q = F.max_pool2d(q, pool, pool).permute(0, 2, 3, 1)
            q = q.reshape(nW, -1, h, hd)
        o = F.scaled_dot_product_attention(q.transpose(1, 2), k.transpose(1, 2), v.transpose(1, 2)).transpose(1, 2)
        wq = wsz // pool
        o = o.reshape(B, Hp // wsz, Wp // wsz, wq, wq, Cc).permute(0, 1, 3, 2, 4, 5).reshape(B, Hp // pool, Wp // pool, Cc)
        o = o[:, :H // pool, :W // pool].reshape(-1, Cc)
        err = (out.float() - o).abs().max().item()
        return err < 3e-2, f"max_abs={err:.3g}"
    if kind == "misc":
        import torch.nn.functional as F
        img = torch.randn(2, 3, 64, 64, device=dev)
        cols = ops.im2col_k7s4(img, 160)
        ref = F.unfold(img, 7, padding=3, stride=4).transpose(1, 2).reshape(-1, 147)
        e1 = (cols[:, :147].float() - ref.to(torch.bfloat16).float()).abs().max().item()
        e1b = cols[:, 147:].float().abs().max().item()
        x = torch.randn(2 * 8 * 8, 40, device=dev)
        mp = ops.maxpool2x2(x, 2, 8, 8)
        refmp = F.max_pool2d(x.view(2, 8, 8, 40).permute(0, 3, 1, 2), 2, 2).permute(0, 2, 3, 1).reshape(-1, 40)
        e2 = (mp - refmp).abs().max().item()
        d = torch.randn(2 * 8 * 8, 24, device=dev); s = torch.randn(2 * 4 * 4, 24, device=dev)
        refd = d.view(2, 8, 8, 24) + F.interpolate(s.view(2, 4, 4, 24).permute(0, 3, 1, 2), scale_factor=2.0, mode="nearest").permute(0, 2, 3, 1)
        ops.add_upsample2x_(d, s, 2, 8, 8)
        e3 = (d.view(2, 8, 8, 24) - refd).abs().max().item()
        t = torch.randn(2 * 50, 37, device=dev)
        add = torch.randn(37, device=dev)
        n = ops.nhwc_to_nchw(t, 2, 50, torch.float32, add)
        e4 = (n - (t.view(2, 50, 37) + add).transpose(1, 2)).abs().max().item()
        back = ops.nchw_to_nhwc(n.contiguous(), torch.float32)
        e5 = (back - (t + add)).abs().max().item()
        ac = ops.add_cast(t, add, torch.float32)
        e6 = (ac - (t + add)).abs().max().item()
        ok = max(e1, e1b, e2, e3, e4, e5, e6) < 1e-5
        return ok, f"im2col={e1:.2g}/{e1b:.2g} maxpool={e2:.2g} upadd={e3:.2g} nchw={e4:.2g} nhwc={e5:.2g} addcast={e6:.2g}"
    raise ValueError(kind)


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "--one":
        kind, kw = sys.argv[2], json.loads(sys.argv[3])
        try:
            ok, msg = run_case(kind, kw)
        except Exception as e:  # noqa
            ok, msg = False, f"EXC {type(e).__name__}: {e}"
        print(("PASS " if ok else "FAIL ") + f"{kind} {kw} :: {msg}", flush=True)
        sys.exit(0 if ok else 1)
    only = sys.argv[1] if len(sys.argv) > 1 else None
    nfail = 0
    for kind, kw in CASES:
        if only and kind != only:
            continue
        try:
            r = subprocess.run([sys.executable, __file__, "--one", kind, json.dumps(kw)],
                               capture_output=True, text=True, timeout=180)
            out = (r.stdout.strip().splitlines() or ["<no output>"])[-1]
            if r.returncode != 0 and not out.startswith("FAIL"):
                out = f"CRASH {kind} {kw} rc={r.returncode} :: {r.stderr.strip()[-600:]}"
            print(out, flush=True)
            nfail += r.returncode != 0
        except subprocess.TimeoutExpired:
            print(f"TIMEOUT {kind} {kw}", flush=True)
            nfail += 1
    print(f"done, failures={nfail}")
