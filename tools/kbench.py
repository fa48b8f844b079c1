#!/usr/bin/env python
"""Micro-timings of single entry points (CUDA events, L2-cold inputs) -- development aid.
    python tools/kbench.py mlp256 [P]
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "project-nerf_b200")]
import torch  # noqa: E402

import b2n  # noqa: E402
from b2n import _lib, ops  # noqa: E402
from src.core import NeuralField  # noqa: E402


def timeit(fn, n=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(n):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    return ts[len(ts) // 2], ts[0]


def mlp256(P):
    torch.manual_seed(0)
    model = NeuralField(dict(mode="part2_nerf", L_embed=10, L_embed_dir=4)).cuda().eval()
    xe = torch.randn(P, 63, device="cuda")
    de = torch.randn(P, 27, device="cuda")
    flops = 2.0 * P * 593408
    for pair in (0, 1):
        b2n._lib.lib.b2n_debug_mlp256_set_pair(pair)
        for save in (False, True):
            med, best = timeit(lambda: ops.nerf_mlp_forward(model.decoder, xe, de, save=save))
            print(f"mlp256 fwd pair={pair} P={P} save={save}: median {med:.3f} ms best {best:.3f} ms -> "
                  f"{flops / best / 1e9:.1f} TFLOP/s ({100 * flops / best / 1e9 / 1391.5:.1f} % of sustained bf16 peak)")
        _, _, _, err = ops.nerf_mlp_forward(model.decoder, xe, de, save=True)
        print(f"  err flag {int(err.item())}")
        # backward chain alone (data gradients): time the autograd call, minus nothing -- report the profiler's entry
        x2 = xe[: P].clone().requires_grad_(True)
        model.train()
        for it in range(6):
            if it == 1:
                b2n._lib.PROFILER = prof = b2n._lib.Profiler()
            r, s_ = model.decoder(x2, de)
            (r.sum() + s_.sum()).backward()
        torch.cuda.synchronize()
        b2n._lib.PROFILER = None
        for name, d in prof.summary().items():
            if "nerf_mlp" in name:
                print(f"    {name}: {d['ms'] / d['calls']:.3f} ms/call")
        model.eval()
    for pair in (0, 1):
        b2n._lib.lib.b2n_debug_mlp256_set_pair(pair)
        for what in ("fwd", "fwd+save", "bwd"):
            prof = torch.zeros(8 + 448, dtype=torch.int64, device="cuda")
            if what == "bwd":
                model.train()
                x2 = xe.clone().requires_grad_(True)
                r, s_ = model.decoder(x2, de)
                torch.cuda.synchronize()
                b2n._lib.lib.b2n_debug_mlp256_flags(16)
                b2n._lib.lib.b2n_debug_mlp256_prof(prof.data_ptr())
                (r.sum() + s_.sum()).backward()
                model.eval()
            else:
                b2n._lib.lib.b2n_debug_mlp256_flags(16)
                b2n._lib.lib.b2n_debug_mlp256_prof(prof.data_ptr())
                ops.nerf_mlp_forward(model.decoder, xe, de, save=what != "fwd")
            torch.cuda.synchronize()
            b2n._lib.lib.b2n_debug_mlp256_prof(None)
            b2n._lib.lib.b2n_debug_mlp256_flags(0)
            pr = prof.tolist()
            n = max(pr[4], 1)
            print(f"pair={pair} {what}: CTA0 cycles per tile pair ({pr[4]} pairs): whole {pr[7] / n:.0f} | pre-step {pr[5] / n:.0f}, "
                  f"bias staging {pr[6] / n:.0f}, epilogue-waits-MMA {pr[2] / n:.0f}, epilogue-body {pr[3] / n:.0f} | "
                  f"MMA-waits-epilogue {pr[0] / n:.0f}, MMA-waits-weights {pr[1] / n:.0f}")
    b2n._lib.lib.b2n_debug_mlp256_set_pair(1)
    with torch.no_grad():
        b2n.set_mlp_precision("fp32")
        med, best = timeit(lambda: model.decoder(xe, de), n=3, warm=1)
        print(f"fp32 layer-wise path: {best:.3f} ms -> {flops / best / 1e9:.1f} TFLOP/s")


def mlp256x(P):
    """role isolation of k_mlp256 (debug flags): kernel time of forward / backward with the epilogue drain and / or the MMAs removed"""
    torch.manual_seed(0)
    model = NeuralField(dict(mode="part2_nerf", L_embed=10, L_embed_dir=4)).cuda().train()
    xe = torch.randn(P, 63, device="cuda")
    de = torch.randn(P, 27, device="cuda")
    lib = b2n._lib.lib
    for pair in (0, 1):
        lib.b2n_debug_mlp256_set_pair(pair)
        for dbg in [int(v) for v in os.environ.get("KB_DBG", "0,1,2,3,7,15").split(",")]:
            lib.b2n_debug_mlp256_flags(dbg)
            x2 = xe.clone().requires_grad_(True)
            for it in range(5):
                if it == 1:
                    b2n._lib.PROFILER = prof = b2n._lib.Profiler()
                r, s_ = model.decoder(x2, de)
                (r.sum() + s_.sum()).backward()
                with torch.no_grad():
                    ops.nerf_mlp_forward(model.decoder, xe, de, save=False)
            torch.cuda.synchronize()
            b2n._lib.PROFILER = None
            lib.b2n_debug_mlp256_flags(0)
            ms = {}
            for name, nbytes, flops, e0, e1 in prof.records:
                if name in ("b2n_nerf_mlp_fwd", "b2n_nerf_mlp_bwd"):
                    key = name[-3:] + ("+save" if nbytes > P * 1000 and name.endswith("fwd") else "")
                    ms.setdefault(key, []).append(e0.elapsed_time(e1))
            print(f"pair={pair} dbg={dbg} ({ {0: 'full', 1: 'no drain', 2: 'no MMA', 3: 'handshakes + weight stream only', 7: 'handshakes only', 15: 'handshakes only, no row loads/stores'}[dbg]}): " +
                  ", ".join(f"{k} {min(v):.3f} ms" for k, v in sorted(ms.items())))
    lib.b2n_debug_mlp256_set_pair(1)


def mlp256t(P):
    """timeline of CTA 0's third tile pair (B2N_TRACE): per step / tile, cycles relative to the pair's first event"""
    torch.manual_seed(0)
    model = NeuralField(dict(mode="part2_nerf", L_embed=10, L_embed_dir=4)).cuda().train()
    xe = torch.randn(P, 63, device="cuda")
    de = torch.randn(P, 27, device="cuda")
    lib = b2n._lib.lib
    names = ["epi:wait", "epi:acc", "epi:done", "mma:wait", "mma:act", "mma:w0", "mma:issued", "epi:stage"]
    for pair, what, cta in ((0, "bwd", 0), (1, "fwd+save", 0), (1, "bwd", 0), (1, "bwd", 1)):
        lib.b2n_debug_mlp256_set_pair(pair)
        lib.b2n_debug_mlp256_flags(32 * cta)
        if True:
            x2 = xe.clone().requires_grad_(True)
            for _ in range(2):
                r, s_ = model.decoder(x2, de)
                (r.sum() + s_.sum()).backward()
            torch.cuda.synchronize()
            prof = torch.zeros(8 + 448, dtype=torch.int64, device="cuda")
            if what == "bwd":
                r, s_ = model.decoder(x2, de)
                torch.cuda.synchronize()
                lib.b2n_debug_mlp256_prof(prof.data_ptr())
                (r.sum() + s_.sum()).backward()
            else:
                lib.b2n_debug_mlp256_prof(prof.data_ptr())
                ops.nerf_mlp_forward(model.decoder, xe, de, save=True)
            torch.cuda.synchronize()
            lib.b2n_debug_mlp256_prof(None)
            tr = prof[8:].view(14, 2, 16)[:, :, :8].cpu()
            t00 = int(tr[tr > 0].min())
            print(f"--- pair={pair} {what} CTA {cta}: cycles since the first event of the pair; columns = {names}")
            for s in range(14):
                for t in range(2):
                    row = tr[s, t]
                    if int(row.max()) == 0:
                        continue
                    print(f"  step {s:2d} tile {t}: " + " ".join(f"{(int(v) - t00) if v > 0 else -1:7d}" for v in row))
    lib.b2n_debug_mlp256_set_pair(1)
    lib.b2n_debug_mlp256_flags(0)


def occ_update():
    """DensityGrid.update: sigma-only sweep through NeuralField.density against the reference-style sweep
    through the full model (zero view directions, rgb dropped), C2 / C4 / C5 shapes, bf16 mode."""
    from src.renderer import DensityGrid

    class NoDensity(torch.nn.Module):
        def __init__(self, m):
            super().__init__()
            self.m, self.mode = m, m.mode

        def forward(self, *a, **k):
            return self.m(*a, **k)

    b2n.set_mlp_precision("bf16")
    cfgs = {
        "C2 part2_instant R=128": (dict(mode="part2_instant", scene_bound=1.5), 128, None),
        "C4 part3 instant R=128": (dict(mode="part3", canonical_type="instant", scene_bound=1.5), 128, 0.3),
        "C5 part4 R=64": (dict(mode="part4", scene_bound=1.5, log2_hashmap_size=20, deform_n_levels=12,
                               deform_log2_hashmap_size=16), 64, None),
    }
    for name, (cfg, R, tval) in cfgs.items():
        torch.manual_seed(0)
        model = NeuralField(cfg).cuda().eval()
        for label, mdl in (("density branch", model), ("full forward", NoDensity(model))):
            grid = DensityGrid(resolution=R, bound=1.5, threshold=0.01).cuda()
            t = None if tval is None else torch.tensor([[tval]], device="cuda")
            p0 = next(model.parameters())

            def once():
                with torch.no_grad():
                    p0.add_(0.0)                 # an optimizer step happened: version counter bumped, sweep cache invalid
                grid.update(mdl, device="cuda", time=t)

            def trio():                          # run.py:1972-1986: three update() calls in a row between two optimizer steps
                with torch.no_grad():
                    p0.add_(0.0)
                for _ in range(3):
                    grid.update(mdl, device="cuda", time=t, decay=0.95)
            med, best = timeit(once, n=5, warm=2)
            med3, best3 = timeit(trio, n=5, warm=2)
            print(f"occ update {name}: {label}: one call median {med:.3f} ms best {best:.3f} ms; "
                  f"three calls in a row (run.py pattern) median {med3:.3f} ms")


def c1_step(P_rays=4096, N=64):
    """C1: vanilla NeRF training step (render_rays + MSE + backward + Adam), B=4096, N=64."""
    from b2n import synthetic
    from src.renderer import render_rays
    for mode in ("bf16", "fp32"):
        b2n.set_mlp_precision(mode)
        torch.manual_seed(0)
        model = NeuralField(dict(mode="part2_nerf", L_embed=10, L_embed_dir=4)).cuda().train()
        opt = torch.optim.Adam(model.parameters(), lr=5e-4)
        ro, rd, tgt = (t.cuda() for t in synthetic.random_rays(P_rays, seed=1))

        def step():
            pred, _, _ = render_rays(model, ro, rd, 2.0, 6.0, N, True, white_bkgd=True)
            loss = torch.nn.functional.mse_loss(pred, tgt[:, :3])
            opt.zero_grad()
            loss.backward()
            opt.step()

        med, best = timeit(step, n=5 if mode == "fp32" else 20, warm=3)
        if mode == "bf16" and "--prof" in sys.argv:
            from torch.profiler import ProfilerActivity, profile
            with profile(activities=[ProfilerActivity.CUDA]) as prof:
                for _ in range(3):
                    step()
                torch.cuda.synchronize()
            print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=25, max_name_column_width=60))
        flops = 3 * 2.0 * P_rays * N * 593408
        print(f"C1 train step [{mode}] B={P_rays} N={N}: median {med:.3f} ms -> {P_rays / med * 1e3 / 1e6:.3f} M rays/s, "
              f"{flops / med / 1e9:.1f} TFLOP/s algorithmic")
        model.eval()
        with torch.no_grad():
            med, best = timeit(lambda: render_rays(model, ro, rd, 2.0, 6.0, N, False, white_bkgd=True), n=10, warm=2)
        print(f"C1 render [{mode}]: median {med:.3f} ms -> {P_rays * N / med * 1e3 / 1e6:.1f} Msamples/s")


def hash_levels(P=24_000_000):
    """hash fwd / bwd time as a function of the level range (contention study)."""
    from b2n import synthetic, march
    torch.manual_seed(0)
    ro, rd, _ = (t.cuda() for t in synthetic.random_rays(2 ** 18, seed=1))
    u = torch.rand(2 ** 18, 128, device="cuda")
    occ = torch.ones(128, 128, 128, dtype=torch.bool, device="cuda")
    m = march.march(ro, rd, 2.0, 6.0, 128, u, bits=march.pack_occupancy(occ), R=128, bound=1.5)
    x = m.pts
    print("points", x.shape[0])
    for (nl, base, note) in ((16, 16, "all 16 levels"), (4, 16, "levels 0-3 (dense 16..54)"), (8, 16, "levels 0-7"),
                             (4, 411, "4 fine hashed levels (res 411..1384)")):
        geom = b2n.HashGeometry(nl, base, 1.5, 19, 2)
        table = torch.randn(geom.n_params, device="cuda", requires_grad=True) * 0.1
        table = table.detach().requires_grad_(True)
        y = b2n.hash_encode(x, table, geom, 1.5)
        g = torch.randn_like(y)
        f_med, _ = timeit(lambda: b2n.hash_encode(x, table, geom, 1.5), n=5, warm=2)

        def bwd():
            y = b2n.hash_encode(x, table, geom, 1.5)
            y.backward(g)
        b_med, _ = timeit(bwd, n=5, warm=2)
        print(f"{note:45s} L={nl:2d}: fwd {f_med:7.3f} ms ({f_med / nl:6.3f}/level)  bwd(+fwd+memset) {b_med - f_med:7.3f} ms "
              f"({(b_med - f_med) / nl:6.3f}/level)")


def hash_sorted():
    """What spatial order buys the hash-grid kernels AS THEY ARE (lanes = the 16 levels of a point): the C2 sample set in
    ray order against the same points sorted by a 30-bit Morton key (torch argsort; sort cost reported separately)."""
    from b2n import synthetic, march
    torch.manual_seed(0)
    ro, rd, _ = (t.cuda() for t in synthetic.random_rays(2 ** 18, seed=1))
    u = torch.rand(2 ** 18, 128, device="cuda")
    occ = torch.ones(128, 128, 128, dtype=torch.bool, device="cuda")
    x = march.march(ro, rd, 2.0, 6.0, 128, u, bits=march.pack_occupancy(occ), R=128, bound=1.5).pts
    print("points", x.shape[0])

    def spread(v):                                   # 10 bits -> every third bit
        v = (v | (v << 16)) & 0x030000FF
        v = (v | (v << 8)) & 0x0300F00F
        v = (v | (v << 4)) & 0x030C30C3
        return (v | (v << 2)) & 0x09249249

    def morton(p):
        q = ((p + 1.5) / 3.0 * 1023.0).clamp(0, 1023).to(torch.int64)
        return spread(q[:, 0]) | (spread(q[:, 1]) << 1) | (spread(q[:, 2]) << 2)

    t_sort, _ = timeit(lambda: torch.argsort(morton(x)), n=3, warm=1)
    order = torch.argsort(morton(x))
    xs = x[order].contiguous()
    print(f"morton key + torch.argsort: {t_sort:.3f} ms (a radix sort of 30-bit keys would be the product path)")
    geom = b2n.HashGeometry(16, 16, 1.5, 19, 2)
    table = (torch.randn(geom.n_params, device="cuda") * 0.1).requires_grad_(True)
    for name, pts in (("ray order", x), ("morton order", xs)):
        y = b2n.hash_encode(pts, table, geom, 1.5)
        g = torch.randn_like(y)
        b2n._lib.PROFILER = prof = b2n._lib.Profiler()
        for _ in range(4):
            y = b2n.hash_encode(pts, table, geom, 1.5)
            y.backward(g)
        torch.cuda.synchronize()
        b2n._lib.PROFILER = None
        print(f"{name:13s}: " + ", ".join(f"{k} {min(e0.elapsed_time(e1) for n_, _, _, e0, e1 in prof.records if n_ == k):.3f} ms"
                                           for k in ("b2n_hash_fwd", "b2n_hash_bwd")))


def hash_variants():
    """A/B of the F = 2 hash-grid kernels on the C2 sample set (24 M active points, L = 16, T = 2^19) and the C5 canonical
    geometry (T = 2^20): (point, level)-per-lane kernels against the pair-lane kernels, and the run-merging threshold."""
    from b2n import synthetic, march
    lib = b2n._lib.lib
    torch.manual_seed(0)
    ro, rd, _ = (t.cuda() for t in synthetic.random_rays(2 ** 18, seed=1))
    u = torch.rand(2 ** 18, 128, device="cuda")
    occ = torch.ones(128, 128, 128, dtype=torch.bool, device="cuda")
    x = march.march(ro, rd, 2.0, 6.0, 128, u, bits=march.pack_occupancy(occ), R=128, bound=1.5).pts
    print("points", x.shape[0])
    for log2T in (19, 20):
        geom = b2n.HashGeometry(16, 16, 1.5, log2T, 2)
        table = (torch.randn(geom.n_params, device="cuda") * 0.1).requires_grad_(True)
        lib.b2n_debug_hash_variant(0, 64)
        y_ref = b2n.hash_encode(x, table, geom, 1.5)
        g = torch.randn_like(y_ref)
        gt_ref, = torch.autograd.grad((y_ref * g).sum(), table)
        for variant, merge in ((0, 64), (1, 64), (2, 0), (2, 24), (2, 64), (2, 128), (2, 200), (3, 64)):
            lib.b2n_debug_hash_variant(variant, merge)
            y = b2n.hash_encode(x, table, geom, 1.5)
            gt, = torch.autograd.grad((y * g).sum(), table)
            ey = float((y - y_ref).abs().max() / y_ref.abs().max())
            eg = float((gt - gt_ref).abs().max() / gt_ref.abs().max())
            b2n._lib.PROFILER = prof = b2n._lib.Profiler()
            for _ in range(4):
                y = b2n.hash_encode(x, table, geom, 1.5)
                y.backward(g)
            torch.cuda.synchronize()
            b2n._lib.PROFILER = None
            print(f"T=2^{log2T} variant={variant} merge_res={merge:3d}: " +
                  ", ".join(f"{k} {min(e0.elapsed_time(e1) for n_, _, _, e0, e1 in prof.records if n_ == k):.3f} ms"
                            for k in ("b2n_hash_fwd", "b2n_hash_bwd")) + f"   max rel diff vs variant 0: y {ey:.2e} g_table {eg:.2e}")
    lib.b2n_debug_hash_variant(3, 64)


def red_bench():
    """red.global.add throughput on an L2-resident table: per lane or per sector?"""
    from b2n._lib import call, ptr, stream
    names = {0: "v2 random lanes", 1: "v2 lane pairs share 16 B", 2: "v2 4 lanes share a sector", 3: "v4 random lanes",
             4: "v4 even lanes only", 5: "v2 even lanes only", 6: "v2 16 lanes share a 128-B line"}
    for log2n in (18, 22):
        n = 1 << log2n
        table = torch.zeros(n, 2, device="cuda")
        blocks, per_thread = 148 * 16, 256
        for mode in range(7):
            fn = lambda: call("b2n_debug_red_bench", ptr(table), n, blocks, per_thread, mode, stream())
            med, best = timeit(fn, n=5, warm=2)
            lanes = blocks * 256 * per_thread * (0.5 if mode in (4, 5) else 1.0)
            print(f"table {n * 8 >> 20:3d} MiB mode {mode} ({names[mode]:32s}): {lanes / best / 1e6:8.1f} G lane-ops/s, "
                  f"{blocks * 8 * per_thread / best / 1e6:7.2f} G warp-instr/s")


def l2_gather():
    """Random 8-byte gathers over tables of 2 MiB .. 512 MiB: the L2 (and beyond-L2) gather peak."""
    from b2n._lib import call, ptr, stream
    sink = torch.zeros(1, device="cuda")
    out = {}
    for log2n in (18, 21, 22, 23, 24, 26):          # float2 entries: 2 MiB, 16, 32, 64, 128, 512 MiB
        n = 1 << log2n
        table = torch.randn(n, 2, device="cuda")
        blocks, per_thread = 148 * 16, 512
        fn = lambda: call("b2n_debug_gather_bench", ptr(table), n, blocks, per_thread, ptr(sink), stream())
        med, best = timeit(fn, n=5, warm=2)
        g = blocks * 256 * per_thread
        out[n * 8 >> 20] = g / best / 1e6
        print(f"table {n * 8 >> 20:4d} MiB: {g / best / 1e6:8.1f} G gathers/s  = {g * 8 / best / 1e6:7.1f} GB/s useful, "
              f"{g * 32 / best / 1e6:7.1f} GB/s of 32-byte sectors")
    return out


def composite(B=2 ** 18, N=128):
    """compositing forward / backward on the C2 shapes: dense occupancy (~72 % of samples active, compact layout)"""
    from b2n import synthetic, march
    torch.manual_seed(0)
    ro, rd, _ = (t.cuda() for t in synthetic.random_rays(B, seed=1))
    u = torch.rand(B, N, device="cuda")
    for kind in ("dense", "sparse"):
        occ = torch.ones(128, 128, 128, dtype=torch.bool, device="cuda") if kind == "dense" else synthetic.ball_occupancy(128, 1.5).cuda()
        m = march.march(ro, rd, 2.0, 6.0, N, u, bits=march.pack_occupancy(occ), R=128, bound=1.5)
        Pn = m.pts.shape[0]
        rgb = torch.rand(Pn, 3, device="cuda", requires_grad=True)
        sigma = (torch.rand(Pn, device="cuda") * 2).requires_grad_(True)
        bg = torch.ones(3, device="cuda")
        from b2n import _lib
        g = torch.randn(B, 3, device="cuda")

        def both():
            c, d, a, _ = b2n.composite(rgb, sigma, m.z, rd, bg=bg, mask_words=m.mask_words, ray_offset=m.ray_offset)
            torch.autograd.grad((c * g).sum(), [rgb, sigma])
        for _ in range(3):
            both()
        torch.cuda.synchronize()
        prof = _lib.Profiler()
        _lib.PROFILER = prof
        for _ in range(20):
            both()
        torch.cuda.synchronize()
        _lib.PROFILER = None
        for name, v in prof.summary().items():
            print(f"composite [{kind}] B={B} N={N} active={Pn}: {name} {v['ms'] / v['calls']:.4f} ms = "
                  f"{v['bytes'] / v['ms'] / 1e6:.0f} GB/s algorithmic ({100 * v['bytes'] / v['ms'] / 1e6 / 6496.8:.1f} % of measured HBM peak)")


def mlp64(P=1 << 22):
    """fused Instant decoder: forward (mma.sync vs tcgen05) and backward, C2-sized (4.2 M points) and Part-3 width"""
    from oracle import nerf_oracle as O
    for pos_dim in (32, 53):
        gen = torch.Generator().manual_seed(0)
        sp = O._fused_init(pos_dim, 16, 64, 1, gen).cuda().requires_grad_(True)
        cp = O._fused_init(43, 3, 64, 2, gen).cuda().requires_grad_(True)
        x = (torch.randn(P, pos_dim, device="cuda") * 0.5).requires_grad_(True)
        d = torch.nn.functional.normalize(torch.randn(P, 3, device="cuda"), dim=-1)
        bands = O.fourier_bands(4).cuda()
        for tc, slots in ((False, 1), (True, 1), (True, 2)):
            ops.INSTANT_FWD_TC = tc
            _lib.lib.b2n_debug_instant_fwd_slots(slots)
            with torch.no_grad():
                med, best = timeit(lambda: b2n.instant_mlp(x, d, bands, sp, cp))
                meds, bests = timeit(lambda: b2n.instant_sigma(x, sp))
            print(f"instant fwd pos_dim={pos_dim} P={P} tc={tc} slots={slots}: {med:.3f} ms (best {best:.3f}); sigma-only {meds:.3f} ms")
        b2n.check_errors()
        rgb, sigma = b2n.instant_mlp(x, d, bands, sp, cp)
        g1, g2 = torch.randn_like(rgb), torch.randn_like(sigma)
        for tc, groups in ((False, 1), (True, 1), (True, 3)):
            ops.INSTANT_BWD_TC = tc
            _lib.lib.b2n_debug_instant_bwd_groups(groups)
            med, best = timeit(lambda: torch.autograd.grad([rgb, sigma], [x, sp, cp], [g1, g2], retain_graph=True))
            print(f"instant bwd pos_dim={pos_dim} P={P} tc={tc} groups={groups}: {med:.3f} ms (best {best:.3f})")
        b2n.check_errors()


def overlap(P=12_000_000):
    """can the L2-atomic-bound table scatter and the SM-bound decoder backward share the SMs?  Independent inputs, two streams."""
    from oracle import nerf_oracle as O
    from src.embeddings import HashGridEncoding
    gen = torch.Generator().manual_seed(0)
    sp = O._fused_init(32, 16, 64, 1, gen).cuda().requires_grad_(True)
    cp = O._fused_init(43, 3, 64, 2, gen).cuda().requires_grad_(True)
    x = (torch.randn(P, 32, device="cuda") * 0.5).requires_grad_(True)
    d = torch.nn.functional.normalize(torch.randn(P, 3, device="cuda"), dim=-1)
    bands = O.fourier_bands(4).cuda()
    # autograd runs a backward node on the stream of its forward: issue each forward on the stream its backward should use
    s1, s2 = torch.cuda.Stream(priority=-1), torch.cuda.Stream()
    torch.cuda.synchronize()
    with torch.cuda.stream(s1):
        rgb, sigma = b2n.instant_mlp(x, d, bands, sp, cp)
        g1, g2 = torch.randn_like(rgb), torch.randn_like(sigma)
    enc = HashGridEncoding(3, dict(otype="HashGrid", n_levels=16, n_features_per_level=2, log2_hashmap_size=19,
                                   base_resolution=16, per_level_scale=1.5)).cuda()
    # ray-ordered samples like the training step: 128 samples along random rays
    B = P // 128
    o = (torch.rand(B, 1, 3, device="cuda") - 0.5) * 2
    dd = torch.nn.functional.normalize(torch.randn(B, 1, 3, device="cuda"), dim=-1)
    tt = torch.linspace(0, 2.0, 128, device="cuda").view(1, 128, 1)
    pts = ((o + dd * tt).reshape(-1, 3).clamp(-1.5, 1.5) / 3.0 + 0.5).contiguous()          # unit cube
    torch.cuda.synchronize()
    with torch.cuda.stream(s2):
        feat = enc(pts)
        gf = torch.randn_like(feat)
    torch.cuda.synchronize()

    def mlp():
        torch.autograd.grad([rgb, sigma], [x, sp, cp], [g1, g2], retain_graph=True)

    def hashb():
        torch.autograd.grad(feat, enc.params, gf, retain_graph=True)

    def both():
        cur = torch.cuda.current_stream()
        s1.wait_stream(cur), s2.wait_stream(cur)
        with torch.cuda.stream(s1):
            mlp()
        with torch.cuda.stream(s2):
            hashb()
        cur.wait_stream(s1), cur.wait_stream(s2)

    for groups in (3, 1, 4):
        _lib.lib.b2n_debug_instant_bwd_groups(groups)
        tm, _ = timeit(mlp)
        th, _ = timeit(hashb)
        tb, _ = timeit(both)
        print(f"groups={groups}: decoder bwd {tm:.3f} ms, table scatter {th:.3f} ms, both on two streams {tb:.3f} ms (sum {tm + th:.3f})")
    _lib.lib.b2n_debug_instant_bwd_groups(3)


if __name__ == "__main__":
    what = sys.argv[1] if len(sys.argv) > 1 else "mlp256"
    P = int(sys.argv[2]) if len(sys.argv) > 2 and sys.argv[2].isdigit() else 262144
    if what == "c1":
        c1_step()
    elif what == "occ":
        occ_update()
    elif what == "hashsort":
        hash_sorted()
    elif what == "hash":
        hash_levels()
    elif what == "hashv":
        hash_variants()
    elif what == "red":
        red_bench()
    elif what == "l2":
        l2_gather()
    elif what == "overlap":
        overlap()
    elif what == "mlp64":
        mlp64()
    elif what == "composite":
        composite()
        composite(8192, 64)
    else:
        {"mlp256": mlp256, "mlp256x": mlp256x, "mlp256t": mlp256t}[what](P)
