#!/usr/bin/env python
"""bench.py -- the CNC hot path on B200, one JSON line (contract in the task brief / DESIGN.md).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--samples N_s]

Workload (BASELINE.json configs[1], restated in SURVEY 8d "Config 2"): the product field layout
(F=8; 3D grid 12 levels res 18..514, T=2^19; 3 planes x 4 levels res 130..1026, T=2^17; MLPs
255->160->80 and 95->160->160->3), random-init weights, +-1-worst-case tables, N_s = 262144 sample
positions per GPU drawn inside the radius-1 ball of the +-1.5 aabb with unit view directions
(synthetic; no datasets offline).  A step = one forward pass (sigma + rgb) of the field over the
batch.  `value` = samples/s with inputs resident in HBM; `e2e` = the same call fed from pinned
host memory with the rgb/sigma result read back, copies inside the timed span.

`--impl reference` runs the same workloads through the UNMODIFIED reference: its own classes
(NGPRadianceField_mygrid_2D3D, OccGridEstimator, render_image_with_occgrid[_test], CNC_context_models; byte-compiled
from /root/reference into oracle/_ref/py) on its own CUDA kernels (compiled unmodified into oracle/_ref/*.so), see
oracle/ref_py.py.  Third-party pieces absent offline stand in as oracle restatements: torchac (the oracle C coder, one
CPU thread like torchac) and tinycudann's SH encoding.  Only the step loop of the training script
(train_CNC_nerf_synthetic.py:302-366) is restated here, because the script is not importable.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

R3 = [18, 24, 33, 44, 59, 80, 108, 148, 201, 275, 376, 514]
R2 = [130, 258, 514, 1026]
F = 8
BYTES_PER_POINT_FWD = 12 + (12 * 8 + 3 * 4 * 4) * 32 + (96 + 96) * 4  # 5388 B/point, SURVEY 8(d)
FLOP_PER_SAMPLE_FWD = 189760


class Clocks:
    """SM clock and throttle-reason sampler running during the timed region (B200_PROFILING.md 'clocks line').
    NVML is polled from a thread every ~2 ms (the timed region of the forward bench is only a few milliseconds, shorter
    than one period of `nvidia-smi -lms`); nvidia-smi is the fallback when pynvml is missing."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc, self.stop, self.t, self.source = index, [], None, False, None, None

    def _handle(self):
        import pynvml as N

        N.nvmlInit()
        try:
            uuid = str(torch.cuda.get_device_properties(self.index).uuid)
            return N, N.nvmlDeviceGetHandleByUUID(("GPU-" + uuid).encode() if not uuid.startswith("GPU-") else uuid.encode())
        except Exception:
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = int(vis.split(",")[self.index]) if vis and vis.split(",")[self.index].isdigit() else self.index
            return N, N.nvmlDeviceGetHandleByIndex(phys)

    def sample_now(self):
        """one synchronous sample (called between the last launch and the synchronize of the timed region: the GPU is
        still executing the queued steps)"""
        if getattr(self, "_nvml", None) is not None:
            self._one(*self._nvml)

    def _one(self, N, h, mx, names, reasons_fn):
        try:
            sm = N.nvmlDeviceGetClockInfo(h, N.NVML_CLOCK_SM)
            r = reasons_fn(h)
            self.rows.append([str(self.index), str(sm), str(mx), "0"] + ["Active" if r & bit else "Not Active" for _, bit in names])
        except Exception:
            pass

    def _poll(self, N, h):
        names = (("hw_slowdown", getattr(N, "nvmlClocksEventReasonHwSlowdown", 0x8)),
                 ("hw_thermal_slowdown", getattr(N, "nvmlClocksEventReasonHwThermalSlowdown", 0x40)),
                 ("sw_thermal_slowdown", getattr(N, "nvmlClocksEventReasonSwThermalSlowdown", 0x20)),
                 ("sw_power_cap", getattr(N, "nvmlClocksEventReasonSwPowerCap", 0x4)))
        reasons_fn = getattr(N, "nvmlDeviceGetCurrentClocksEventReasons", None) or N.nvmlDeviceGetCurrentClocksThrottleReasons
        mx = N.nvmlDeviceGetMaxClockInfo(h, N.NVML_CLOCK_SM)
        self._nvml = (N, h, mx, names, reasons_fn)
        while not self.stop:
            self._one(N, h, mx, names, reasons_fn)
            time.sleep(0.002)

    def __enter__(self):
        try:
            N, h = self._handle()
            self.source = "nvml"
            self.t = threading.Thread(target=self._poll, args=(N, h), daemon=True)
            self.t.start()
            return self
        except Exception:
            pass
        try:
            self.source = "nvidia-smi"
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-i", str(self.index), "-lms", "100"], stdout=subprocess.PIPE, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def __exit__(self, *a):
        self.stop = True
        if self.proc:
            time.sleep(0.12)
            self.proc.terminate()
        if self.t:
            self.t.join(timeout=2)

    def summary(self):
        sm = [float(r[1]) for r in self.rows if len(r) >= 8 and r[1].replace(".", "").isdigit()]
        mx = [float(r[2]) for r in self.rows if len(r) >= 8 and r[2].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({n for r in self.rows if len(r) >= 8 for n, v in zip(names, r[4:8]) if v.startswith("Active")})
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm), "source": self.source}


def peaks():
    try:
        p = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        return float(p["hbm_gbs"]), "measured"
    except Exception:
        return 6650.0, "fallback"


def ncu_traffic(kernel, n_samples):
    """dram__bytes_read.sum + dram__bytes_write.sum of one launch, from the committed `ncu --set full` summary
    (profiles/ncu_traffic.json, written by scripts/ncu_summary.py); None when no capture matches this launch size"""
    try:
        t = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json")))
        e = t.get(f"{kernel}@{n_samples}")
        return None if e is None else int(e["dram_bytes"])
    except Exception:
        return None


def make_inputs(n, seed, device):
    g = torch.Generator(device="cpu").manual_seed(seed)
    d = torch.randn(n, 3, generator=g)
    r = torch.rand(n, 1, generator=g) ** (1 / 3)
    pos = d / d.norm(dim=-1, keepdim=True) * r  # uniform in the radius-1 ball (aabb +-1.5)
    dirs = torch.randn(n, 3, generator=g)
    dirs = dirs / dirs.norm(dim=-1, keepdim=True)
    return pos.float().to(device), dirs.float().to(device)


class Arm:
    """what a leg of the bench needs, from this package (`ours`) or from the unmodified reference (`reference`)"""

    def __init__(self, impl, dev):
        self.impl, self.dev = impl, dev
        if impl == "ours":
            from cnc_b200 import nerfacc, render
            from cnc_b200.context_models import CNC_context_models
            from cnc_b200.field import NGPRadianceField_mygrid_2D3D
            from cnc_b200.gridencoder import GridEncoder

            self.Field, self.GridEncoder, self.CM = NGPRadianceField_mygrid_2D3D, GridEncoder, CNC_context_models
            self.Estimator, self.Rays = nerfacc.OccGridEstimator, render.Rays
            self.render_train, self.render_test = render.render_image_with_occgrid, render.render_image_with_occgrid_test
            self.cm_kw = dict(Rb=128, device=dev)
        else:
            from oracle import ref_py

            r = ref_py.load()
            import datasets.utils as du   # the reference's (resolved by oracle/ref_py)

            self.Field, self.GridEncoder, self.CM = r.ngp.NGPRadianceField_mygrid_2D3D, r.ngp.GridEncoder, r.bpp.CNC_context_models
            self.Estimator, self.Rays = r.nerfacc.OccGridEstimator, du.Rays
            self.render_train, self.render_test = r.utils.render_image_with_occgrid, r.utils.render_image_with_occgrid_test
            self.cm_kw = {}

    def field(self, seed=0, F=F):
        torch.manual_seed(seed)
        f = self.Field(aabb=[-1.5, -1.5, -1.5, 1.5, 1.5, 1.5], n_features_per_level=F, n_neurons=160, resolutions_list=R3,
                       log2_hashmap_size=19, resolutions_list_2D=R2, log2_hashmap_size_2D=17, ste_binary=True).to(self.dev)
        g = torch.Generator(device="cpu").manual_seed(seed + 1)
        with torch.no_grad():  # +-1 with p = 0.5 after STE: worst-case entropy / locality; identical in both arms
            for k in ("xyz", "xy", "xz", "yz"):
                p = getattr(f.mlp_base, f"encoding_{k}").params
                p.copy_(torch.where(torch.rand(p.shape, generator=g) < 0.5, -0.5, 0.5).to(self.dev))
        return f

    def context_model(self, **kw):
        cm = self.CM(num_dim=3, resolutions_list=R3, resolutions_list_2D=R2, log2_hashmap_size=19, log2_hashmap_size_2D=17,
                     n_features=F, sample_num=150000, max_context_layer_num=3, ste_binary=True, skip_levels_3D=(0, 1, 2),
                     skip_levels_2D=(0,), **self.cm_kw, **kw)
        return cm.to(self.dev)

    def estimator(self):
        est = self.Estimator(roi_aabb=[-1.5, -1.5, -1.5, 1.5, 1.5, 1.5], resolution=128, levels=1).to(self.dev)
        c = (torch.arange(128, device=self.dev) + 0.5) / 128 * 3 - 1.5
        X, Y, Z = torch.meshgrid(c, c, c, indexing="ij")
        est.binaries = (X * X + Y * Y + Z * Z <= 1.0).unsqueeze(0)   # radius-1 ball in the +-1.5 aabb (15.5 % of the cells)
        est.occs = est.binaries.reshape(-1).float()
        return est


class RefTrainStep:
    """train_CNC_nerf_synthetic.py:302-366 around the reference's own objects: render_image_with_occgrid -> MSE
    (+ lmbda * bits per parameter) -> GradScaler-scaled backward -> Adam (field) and Adam (context models), as the script
    does it (two optimizers, torch's stock Adam, `scaler.scale(loss).backward()` without unscaling, :362-364)."""

    def __init__(self, arm, field, est, cm=None, lmbda=0.0, lr=1e-4):
        self.arm, self.field, self.est, self.cm, self.lmbda = arm, field, est, cm, lmbda
        self.opt = torch.optim.Adam([{"params": field.parameters()}], lr=lr, eps=1e-15, weight_decay=2e-6)
        self.opt2 = None if cm is None else torch.optim.Adam([{"params": cm.parameters()}], lr=lr, eps=1e-15)
        self.step_id = 0

    def __call__(self, rays, pixels, refresh_occupancy=False):
        self.field.train()
        self.est.train()
        if refresh_occupancy:
            self.est.update_every_n_steps(step=self.step_id, occ_thre=1e-2,
                                          occ_eval_fn=lambda x: self.field.query_density(x) * 5e-3)
        rgb, acc, depth, n = self.arm.render_train(self.field, self.est, rays, render_step_size=5e-3,
                                                   render_bkgd=torch.ones(3, device=pixels.device))
        loss = torch.nn.functional.mse_loss(rgb, pixels)
        if self.cm is not None and self.lmbda > 0:
            mb = self.field.mlp_base
            bpp, _ = self.cm.forward_binary_vxl_mixPg_3D2D(mb.encoding_xyz, mb.encoding_xy, mb.encoding_xz, mb.encoding_yz,
                                                           self.est.binaries, step=self.step_id)
            loss = loss + self.lmbda * bpp
            self.opt2.zero_grad()
        self.opt.zero_grad()
        (loss * 1024.0).backward()
        self.opt.step()
        if self.opt2 is not None:
            self.opt2.step()
        self.step_id += 1
        return loss.detach(), n


def cpu_baseline(n=4096):
    """oracle port (scalar C + numpy MLP) on one host core over a bounded sample of the workload."""
    from oracle import oracle as o

    rng = np.random.default_rng(0)
    x = rng.random((n, 3), dtype=np.float32)
    offs3, offs2 = o.grid_layout(3, R3, 19), o.grid_layout(2, R2, 17)
    t3 = np.where(rng.random((offs3[-1], F)) < 0.5, -1, 1).astype(np.float32)
    t2 = [np.where(rng.random((offs2[-1], F)) < 0.5, -1, 1).astype(np.float32) for _ in range(3)]
    W = [rng.normal(size=s).astype(np.float32) * 0.05 for s in ((255, 160), (160, 80), (95, 160), (160, 160), (160, 3))]
    reps, t0 = 0, time.perf_counter()
    while time.perf_counter() - t0 < 10.0:
        f3 = o.grid_encode_fwd(x, t3, offs3, R3, 12).transpose(1, 0, 2).reshape(n, -1)
        fs = [o.grid_encode_fwd(np.ascontiguousarray(x[:, ax]), t, offs2, R2, 4).transpose(1, 0, 2).reshape(n, -1)
              for ax, t in zip(([0, 1], [0, 2], [1, 2]), t2)]
        h = np.concatenate([f3, *fs, o.freq_embed(x)], 1)
        h = np.maximum(h @ W[0], 0) @ W[1]
        g = np.concatenate([o.sh16(x), h[:, 1:]], 1)
        _ = np.maximum(np.maximum(g @ W[2], 0) @ W[3], 0) @ W[4]
        reps += 1
    dt = time.perf_counter() - t0
    return {"value": reps * n / dt, "unit": "samples/s", "cores": 1, "kind": "port",
            "sample": f"{reps} x {n} samples of the same layout through oracle/ (C gather + numpy fp32 MLP)"}


def codec_bench(arm, cpu_seconds=12.0, rank=0, world=1, dist=None):
    """BASELINE metric (ii): context-model entropy encode + decode of the product hash tables (configs[2]):
    L=12 3D levels (res 18..514, T=2^19) + 3 planes x 4 levels (T=2^17), F=8, ball occupancy, biased +-1 tables,
    random-init context models.  MB/s = fp32 table bytes represented (161.3 MB) / wall time of the public
    encode_/decode_binary_vxl_mixPg_3D2D call (probabilities + range coding + streams on the host / in files).
    Reference arm: the reference's own CNC_context_models on its own kernels, coding each stream with the torchac
    stand-in (oracle C coder) on one CPU thread after a D2H copy, exactly the structure of utils_bpp_acc.py:77-110.
    CPU baseline (ours arm): the bare coder over the same (cdf, symbol) streams on one thread and on all host cores.
    N > 1: the codec of one scene does not shard (a level's largest stream is one serial chain and levels decode in
    order), so every rank codes its own scene -- replicas only; MB/s = N x table bytes / max-over-ranks time."""
    import tempfile

    dev, ours = arm.dev, arm.impl == "ours"
    g = torch.Generator(device="cpu").manual_seed(100 + rank)
    t0 = time.perf_counter()
    encs = [arm.GridEncoder(num_dim=3, n_features=F, resolutions_list=R3, log2_hashmap_size=19, ste_binary=True).to(dev)] + \
           [arm.GridEncoder(num_dim=2, n_features=F, resolutions_list=R2, log2_hashmap_size=17, ste_binary=True).to(dev) for _ in range(3)]
    with torch.no_grad():
        for e in encs:
            e.params.copy_(torch.where(torch.rand(e.params.shape, generator=g) < 0.7, 0.5, -0.5).to(dev))
    torch.cuda.synchronize()
    t1 = time.perf_counter()
    torch.manual_seed(rank)
    cm = arm.context_model()
    torch.cuda.synchronize()
    t_setup = time.perf_counter() - t1
    with torch.no_grad():
        cm.context_model_3D[4].bias.fill_(0.6)
        for sq in cm.context_model_2D:
            sq[0].bias.fill_(0.6)
    vxl = arm.estimator().binaries
    tmp = tempfile.mkdtemp(prefix="cnc_bench_")
    prefix = os.path.join(tmp, "bench")
    captured = {"c1": [], "sym": []}
    ctx_ms = []          # (launch ms, voxels, entries) of every cnc_context3d_probs launch of the timed encodes
    if ours:
        from cnc_b200 import torchac as tac

        orig, orig_fused = tac.encode_streams_async, cm._probs_3D_fused

        def spy(c1s, syms):
            captured["c1"] += list(c1s)
            captured["sym"] += list(syms)
            return orig(c1s, syms)

        def timed_fused(Enc, table, bvx, n, lo, hi, Pg_n):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            out = orig_fused(Enc, table, bvx, n, lo, hi, Pg_n)
            e1.record()
            ctx_ms.append((e0, e1, cm._cs_host(n, hi) - cm._cs_host(n, lo), hi - lo))
            return out

        tac.encode_streams_async = spy
        try:
            cm.encode_binary_vxl_mixPg_3D2D(*encs, vxl, prefix, return_streams=True)  # warm-up (+ capture the streams)
        finally:
            tac.encode_streams_async = orig

    def encode():
        with torch.no_grad():
            if ours:
                return cm.encode_binary_vxl_mixPg_3D2D(*encs, vxl, prefix, return_streams=True)
            return cm.encode_binary_vxl_mixPg_3D2D(*encs, vxl, prefix) + (None,)

    def decode(Pgs, streams):
        recs = [torch.ones_like(e.params) for e in encs]
        with torch.no_grad():
            if ours:
                return cm.decode_binary_vxl_mixPg_3D2D(*encs, *recs, vxl, Pgs, prefix, streams=streams)
            return cm.decode_binary_vxl_mixPg_3D2D(*encs, *recs, vxl, Pgs, prefix)

    def sync():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
            torch.cuda.synchronize()

    if not ours:
        encode()   # warm-up
    enc_s, dec_s = [], []
    for it in range(3 if ours else 2):
        if ours and it == 2:
            cm._probs_3D_fused = timed_fused    # the per-launch events go into the last repetition only
        sync(); t = time.perf_counter()
        Pgs, est_MB, coded_MB, streams = encode()
        sync(); enc_s.append(time.perf_counter() - t)   # after the barrier: the slowest rank's time
        if ours:
            cm._probs_3D_fused = orig_fused
        sync(); t = time.perf_counter()
        out = decode(Pgs, streams)
        sync(); dec_s.append(time.perf_counter() - t)
    ok = all(bool(((torch.where(e.params >= 0, 1.0, -1.0) == r) | (r == 1)).all()) for e, r in zip(encs, out))
    n_params = sum(e.params.numel() for e in encs)
    table_MB = n_params * 4 / 1e6
    if ours:
        n_sym = sum(int(x.numel()) for x in captured["sym"])
    else:   # 8 symbols per coded byte-stream row: count from the tables (every coded row is restored, the others stay +1)
        n_sym = None
    import shutil

    shutil.rmtree(tmp, ignore_errors=True)
    if dist is not None:
        flag = torch.tensor([1.0 if ok else 0.0, min(enc_s), min(dec_s)], dtype=torch.float64, device=dev)
        mn = flag.clone(); dist.all_reduce(mn, op=dist.ReduceOp.MIN)
        mx = flag.clone(); dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        ok = bool(mn[0] > 0)
        enc_s, dec_s = [float(mx[1])], [float(mx[2])]
    if rank != 0:
        return None
    enc, dec = min(enc_s), min(dec_s)
    res = {"workload": "configs[2]: product tables L=12 T=2^19 + 3x4 planes T=2^17, F=8, ball occupancy, 33 streams"
                       + (f"; {world} scenes, one per GPU (replicas only)" if world > 1 else ""),
           "table_MB_fp32": table_MB, "table_MB_1bit": n_params / 8 / 1e6, "coded_MiB": coded_MB, "estimated_MiB": est_MB,
           "encode_s": enc, "decode_s": dec, "encode_MBps": world * table_MB / enc, "decode_MBps": world * table_MB / dec,
           "roundtrip_ok": ok, "scaling": "weak (replicas)" if world > 1 else "n/a", "setup_s": t_setup,
           "setup_what": "CNC_context_models.__init__: inverse hash tables of all levels (utils_bpp_acc.py:294-335)"}
    if not ours:
        res["coder"] = "torchac stand-in: oracle C range coder, one CPU thread, after a D2H copy (utils_bpp_acc.py:77-110)"
        return res
    from oracle import oracle as o

    res.update({"symbols": n_sym, "encode_Msym_per_s": world * n_sym / enc / 1e6, "decode_Msym_per_s": world * n_sym / dec / 1e6,
                "table_state_MB": cm.table_bytes() / 1e6})
    # the same codec from occupancy-pruned tables (histogram pass at construction, vertex lists per occupancy grid)
    if world == 1:
        state = cm.state_dict()
        del cm
        torch.cuda.empty_cache()
        torch.cuda.synchronize()
        t = time.perf_counter()
        torch.manual_seed(rank)
        cm = arm.context_model(tables="pruned")
        torch.cuda.synchronize()
        t_setup_p = time.perf_counter() - t
        cm.load_state_dict(state)
        t = time.perf_counter()
        Pgs, _, coded_p, streams = encode()          # first call: builds the pruned vertex lists of the nine coded levels
        torch.cuda.synchronize()
        t_first = time.perf_counter() - t
        es, ds = [], []
        for _ in range(2):
            sync(); t = time.perf_counter()
            Pgs, _, coded_p, streams = encode()      # (encode() drops the per-call cache: the lists are rebuilt every time)
            sync(); es.append(time.perf_counter() - t)
            sync(); t = time.perf_counter()
            out = decode(Pgs, streams)
            sync(); ds.append(time.perf_counter() - t)
        okp = all(bool(((torch.where(e.params >= 0, 1.0, -1.0) == r) | (r == 1)).all()) for e, r in zip(encs, out))
        res["pruned_tables"] = {"what": "tables='pruned': row statistics from one histogram pass per level at construction; vertex lists "
                                        "built per occupancy grid for the vertices that pass the occupancy test (csrc/table_build.cu)",
                                "setup_s": t_setup_p, "first_encode_s": t_first, "encode_s": min(es), "decode_s": min(ds),
                                "table_state_MB": cm.table_bytes() / 1e6, "coded_MiB": coded_p, "roundtrip_ok": okp}
    # dominant kernel of the encode: cnc_context3d_probs (one launch per coded chunk), timed live with CUDA events
    if ctx_ms:
        hbm_peak, which = peaks()
        ms_l = [e0.elapsed_time(e1) for e0, e1, _, _ in ctx_ms]
        vox, ent = sum(v for _, _, v, _ in ctx_ms), sum(e for _, _, _, e in ctx_ms)
        # algorithmic bytes (codec): per voxel 6 B of int16 coordinates + the masked 3-level context gather
        # 3 x 8 corners x F x 4 B; per entry 8 B of run offsets + F x 4 B of probabilities + 1 B exists flag
        bytes_alg = vox * (6 + 3 * 8 * F * 4) + ent * (8 + 4 * F + 1)
        ach = bytes_alg / (sum(ms_l) * 1e-3) / 1e9
        res["roofline"] = {"bound": "hbm", "kernel": "cnc::context3d_kernel (mask + 3-level masked gather + 25-32-32-8 MLP + overlap-weighted mean)",
                           "achieved": ach, "peak": hbm_peak, "peak_source": which, "unit": "GB/s", "frac": ach / hbm_peak,
                           "traffic": ncu_traffic("context3d_kernel", 524288), "traffic_what": "one launch (a 524 288-entry chunk, profiles/r01_v11_ncu_context3d.txt)", "launches": len(ms_l),
                           "ms_all_launches": sum(ms_l), "voxels": vox, "entries": ent, "algorithmic_bytes": bytes_alg,
                           "note": "the context tables are read as 1-bit planes from L2, so DRAM traffic << algorithmic bytes; the "
                                   "kernel is FFMA-issue bound (2080 FMA per voxel), see profiles/"}
    # CPU coder on a bounded sample: streams in descending size until the time budget is used
    order = sorted(range(len(captured["sym"])), key=lambda k: -captured["sym"][k].numel())
    host = [(captured["c1"][k].cpu().numpy().view(np.uint16), captured["sym"][k].cpu().numpy()) for k in order]
    done_sym, t_cpu = 0, 0.0
    for c1, sy in host:
        t = time.perf_counter()
        data = o.ac_encode(c1, sy)
        o.ac_decode(c1, data)
        t_cpu += time.perf_counter() - t
        done_sym += sy.size
        if t_cpu > cpu_seconds:
            break
    cpu_sym_per_s = done_sym / t_cpu            # encode + decode of each symbol, one thread
    # all host cores: the 33 streams dealt to a thread pool, largest first (ctypes releases the GIL inside the C coder)
    from concurrent.futures import ThreadPoolExecutor

    cores = os.cpu_count() or 1

    def one(cs):
        d = o.ac_encode(cs[0], cs[1])
        o.ac_decode(cs[0], d)
        return cs[1].size

    t = time.perf_counter()
    with ThreadPoolExecutor(max_workers=cores) as ex:
        done_all = sum(ex.map(one, host))
    t_all = time.perf_counter() - t
    res["cpu_baseline"] = {"kind": "port", "cores": 1, "unit": "MB/s", "value": table_MB / (n_sym / cpu_sym_per_s),
                           "Msym_per_s_enc_plus_dec": cpu_sym_per_s / 1e6,
                           "sample": f"{done_sym} of {n_sym} symbols (largest streams first) through oracle/ C range coder, "
                                     "encode+decode, one thread; MB/s = table bytes / (coder time for all symbols, "
                                     "encode+decode); probabilities not included (the reference computes them on the GPU)",
                           "all_cores": {"cores": cores, "value": table_MB / t_all, "unit": "MB/s", "seconds": t_all,
                                         "sample": f"all {len(host)} streams ({done_all} symbols), encode+decode, one thread-pool task per "
                                                   f"stream over {cores} threads: bounded below by the longest stream's serial chain"}}
    return res


SAMPLES_PER_RANK = 380625     # what the 1100 rays of the single-GPU batch (seed 7) produce on the ball scene


def train_bench(arm, rank, world, steps, field):
    """forward + backward + (N > 1: NCCL exchange of the gradients) + Adam on a synthetic ray batch per rank:
    rays from a radius-4 sphere towards the origin, ball occupancy, random target pixels.
    Reference arm: the same step around the reference's own objects (RefTrainStep), single GPU."""
    dev, ours = arm.dev, arm.impl == "ours"
    est = arm.estimator()
    g = torch.Generator(device="cpu").manual_seed(7 + rank)
    n_rays, pool = 1100, 1400
    o = torch.randn(pool, 3, generator=g)
    o = o / o.norm(dim=-1, keepdim=True) * 4
    tgt = (torch.rand(pool, 3, generator=g) - 0.5) * 1.2
    d = tgt - o
    d = d / d.norm(dim=-1, keepdim=True)
    o_all, d_all, px_all = o.to(dev), d.to(dev), torch.rand(pool, 3, generator=g).to(dev)
    rays = arm.Rays(o_all[:n_rays].contiguous(), d_all[:n_rays].contiguous())
    pixels = px_all[:n_rays].contiguous()
    bk = torch.ones(3, device=dev)
    if ours:
        from cnc_b200.trainer import TrainStep

        ts = TrainStep(field, est, lr=1e-4, exchange=os.environ.get("CNC_EXCHANGE", "auto"))
        # the batch of the next step is known while this one runs (here: the same rays), so its occupancy march overlaps this
        # step's forward / backward on a side stream (trainer.TrainStep `next_rays`); `no_lookahead_ms` below is without
        call = lambda t, ahead=True: t(rays, pixels, render_bkgd=bk, refresh_occupancy=False, next_rays=(lambda n: rays) if ahead else None)
    else:
        ts = RefTrainStep(arm, field, est, lr=1e-4)
        call = lambda t: t(rays, pixels, refresh_occupancy=False)
    n_s = 0
    for i in range(4 if world > 1 else 3):
        _, n_s = call(ts)
        if i == 0 and world > 1 and ours:
            # every rank sizes its batch to the sample budget of the single-GPU run, as the training loop sizes each batch from
            # the previous one's sample count (train...:340-344): ranks march different rays, equal work needs unequal ray counts
            n_rays = max(64, min(pool, int(round(n_rays * SAMPLES_PER_RANK / max(int(n_s), 1)))))
            rays = arm.Rays(o_all[:n_rays].contiguous(), d_all[:n_rays].contiguous())
            pixels = px_all[:n_rays].contiguous()
    torch.cuda.synchronize()
    if world > 1:
        torch.distributed.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    tot = 0
    for _ in range(steps):
        _, n_s = call(ts)
        tot += int(n_s)
    e1.record()
    torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1), float(tot)], dtype=torch.float64, device=dev)
    if world > 1:
        ms = t[:1].clone()
        torch.distributed.all_reduce(ms, op=torch.distributed.ReduceOp.MAX)
        cnt = t[1:].clone()
        torch.distributed.all_reduce(cnt, op=torch.distributed.ReduceOp.SUM)
        t = torch.cat([ms, cnt])
    ms, tot = t.tolist()
    out = {"what": ("occupancy march + visibility pass (fused density kernel) + fused forward (cnc_field_fwd_train) + volume "
                    "rendering + MSE + backward (cnc_dgrad / cnc_wgrad / K2) + gradient exchange + fused Adam") if ours else
                   ("reference: estimator.sampling (traverse_grids 2 passes + query_density) + rendering(radiance_field) + MSE + "
                    "autograd backward + torch Adam (train_CNC_nerf_synthetic.py:302-366)"),
           "steps": steps, "ms_per_step": ms / steps, "scaling": "weak",
           "samples_per_step_all_ranks": tot / steps, "samples_per_s": tot / (ms * 1e-3),
           "comm_bytes_per_step": ts.comm_bytes_per_step() if (ours and world > 1) else 0, "rays_per_rank": n_rays}
    if ours:
        for _ in range(2):
            call(ts, False)
        torch.cuda.synchronize()
        if world > 1:
            torch.distributed.barrier()
        e0.record()
        for _ in range(steps):
            call(ts, False)
        e1.record()
        torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
        if world > 1:
            torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
        out["no_lookahead_ms"] = float(t[0]) / steps
        out["lookahead"] = "occupancy march of the next batch issued on a side stream during this step (same samples)"
    if ours and world > 1:
        # the same step without data parallelism (every rank its own replica, no exchange), same run, same rays: the
        # scaling efficiency of the step that has the collective, from one launch
        field1 = arm.field(seed=0)
        ts1 = TrainStep(field1, arm.estimator(), lr=1e-4, data_parallel=False)
        for _ in range(3):
            call(ts1)
        torch.cuda.synchronize()
        torch.distributed.barrier()
        e0.record()
        for _ in range(steps):
            call(ts1)
        e1.record()
        torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
        torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
        out["replicas_without_exchange_ms"] = float(t[0]) / steps
        out["efficiency_vs_replicas"] = out["replicas_without_exchange_ms"] / out["ms_per_step"]
        del ts1, field1
        out["comm"] = ts.comm_description()
        if ts.table_opt is not None and ts.table_opt.peer is not None:
            out["nvlink_bytes_per_step"] = ts.table_opt.link_bytes_per_step()
    if world > 1 and ours:
        # the lambda > 0 step under data parallelism: rays sharded as above, the rate term shared among the ranks
        # (context_models.set_data_parallel: every rank samples 150 000 / N entries and evaluates its share of the plane terms)
        cm = arm.context_model()
        ts2 = TrainStep(field, est, context_model=cm, lmbda=1e-3, lr=1e-4, exchange=os.environ.get("CNC_EXCHANGE", "auto"))
        for _ in range(2):
            call(ts2)
        torch.cuda.synchronize()
        torch.distributed.barrier()
        e0.record()
        for _ in range(steps):
            call(ts2)
        e1.record()
        torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
        torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
        out["with_rate_term"] = {"what": "lambda > 0 step, data parallel: rate term shared among the ranks (sampled entries and "
                                         "plane terms split, gradients averaged)", "lambda": 1e-3, "ms_per_step": float(t[0]) / steps,
                                 "sampled_entries_per_rank": int(cm._dp_snl[0].sum()) if getattr(cm, "_dp_snl", None) else None}
        del cm, ts2
        torch.cuda.empty_cache()
    if world == 1:
        # the same step with the rate term of the CNC loss (lambda > 0: context model on 150 000 sampled entries + planes)
        cm = arm.context_model()
        if ours:
            ts2 = TrainStep(field, est, context_model=cm, lmbda=1e-3, lr=1e-4, exchange=os.environ.get("CNC_EXCHANGE", "auto"))
        else:
            ts2 = RefTrainStep(arm, field, est, cm=cm, lmbda=1e-3, lr=1e-4)
        for _ in range(2):
            call(ts2)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(steps):
            call(ts2)
        e1.record()
        torch.cuda.synchronize()
        out["with_rate_term"] = {"what": "same step + lambda * bits-per-parameter (forward_binary_vxl_mixPg_3D2D, 150 000 sampled "
                                         "entries, dimension-wise context) and its backward", "lambda": 1e-3,
                                 "ms_per_step": e0.elapsed_time(e1) / steps}
        if ours:     # the same with the rate term on a side stream beside the render path (TrainStep.overlap_rate_term)
            ts2.overlap_rate_term = True
            call(ts2)
            torch.cuda.synchronize()
            e0.record()
            for _ in range(steps):
                call(ts2)
            e1.record()
            torch.cuda.synchronize()
            out["with_rate_term"]["two_streams_ms"] = e0.elapsed_time(e1) / steps
            ts2.overlap_rate_term = False
        # the rate term alone (forward + backward), the part of the step utils_bpp_acc.py:533-706 owns
        mb = field.mlp_base
        encs = (mb.encoding_xyz, mb.encoding_xy, mb.encoding_xz, mb.encoding_yz)

        def rate_only(k):
            bpp, _ = cm.forward_binary_vxl_mixPg_3D2D(*encs, est.binaries, step=k)
            bpp.backward()
            for p in list(field.parameters()) + list(cm.parameters()):
                p.grad = None

        rate_only(1)
        torch.cuda.synchronize()
        e0.record()
        for k in range(steps):
            rate_only(1 + k)
        e1.record()
        torch.cuda.synchronize()
        out["with_rate_term"]["rate_term_alone_ms"] = e0.elapsed_time(e1) / steps
        del cm, ts2
        torch.cuda.empty_cache()
        # test-time renderer: one 256 x 256 view of the same scene through render_image_with_occgrid_test
        H = W = 256
        v, u = torch.meshgrid(torch.linspace(-0.3, 0.3, H, device=dev), torch.linspace(-0.3, 0.3, W, device=dev), indexing="ij")
        dirs = torch.stack([u, v, torch.ones_like(u)], -1)
        dirs = dirs / dirs.norm(dim=-1, keepdim=True)
        img = arm.Rays(torch.tensor([0.0, 0.0, -4.0], device=dev).expand(H, W, 3).contiguous(), dirs.contiguous())
        was_training = field.training
        field.eval()
        kw = dict(render_step_size=5e-3, render_bkgd=bk)
        arm.render_test(1024, field, est, img, **kw)
        torch.cuda.synchronize()
        e0.record()
        _, opa, _, n_tot = arm.render_test(1024, field, est, img, **kw)
        e1.record()
        torch.cuda.synchronize()
        ms_img = e0.elapsed_time(e1)
        out["test_render"] = {"what": "render_image_with_occgrid_test (examples/utils.py:316-489), the reference's round schedule, "
                                      "early stop at 1 - 1e-4, untrained field (low density: most rays run to the far side)",
                              "rays": H * W, "samples": int(n_tot), "ms_per_image": ms_img,
                              "rays_per_s": H * W / (ms_img * 1e-3), "samples_per_s": n_tot / (ms_img * 1e-3),
                              "mean_opacity": float(opa.mean())}
        # the evaluation size of the training scripts: one 800 x 800 view (640 000 rays), same camera
        H8 = W8 = 800
        v8, u8 = torch.meshgrid(torch.linspace(-0.3, 0.3, H8, device=dev), torch.linspace(-0.3, 0.3, W8, device=dev), indexing="ij")
        d8 = torch.stack([u8, v8, torch.ones_like(u8)], -1)
        d8 = d8 / d8.norm(dim=-1, keepdim=True)
        img8 = arm.Rays(torch.tensor([0.0, 0.0, -4.0], device=dev).expand(H8, W8, 3).contiguous(), d8.contiguous())
        if ours:
            arm.render_test(1024, field, est, img8, **kw)   # warm-up (the reference arm is timed cold: one view is 5 s)
        torch.cuda.synchronize()
        e0.record()
        _, _, _, n8 = arm.render_test(1024, field, est, img8, **kw)
        e1.record()
        torch.cuda.synchronize()
        out["test_render"]["view_800x800"] = {"rays": H8 * W8, "samples": int(n8), "ms_per_image": e0.elapsed_time(e1),
                                              "samples_per_s": n8 / (e0.elapsed_time(e1) * 1e-3)}
        del img8, d8, u8, v8
        if ours:
            arm.render_test(1024, field, est, img, device_loop=False, **kw)
            torch.cuda.synchronize()
            e0.record()
            arm.render_test(1024, field, est, img, device_loop=False, **kw)
            e1.record()
            torch.cuda.synchronize()
            out["test_render"]["host_loop_ms_per_image"] = e0.elapsed_time(e1)
            out["test_render"]["what"] += "; sync-free device loop (wf_* kernels + cnc_field_fwd_n), host_loop_ms = the python loop on the same kernels"
            arm.render_test(1024, field, est, img, samples_per_round=32, **kw)
            torch.cuda.synchronize()
            e0.record()
            _, _, _, n_k = arm.render_test(1024, field, est, img, samples_per_round=32, **kw)
            e1.record()
            torch.cuda.synchronize()
            out["test_render"]["fixed_rounds_of_32"] = {"what": "same view, samples_per_round=32 (fewer, larger rounds; same image within "
                                                               "early_stop_eps)", "samples": int(n_k), "ms_per_image": e0.elapsed_time(e1),
                                                       "rays_per_s": H * W / (e0.elapsed_time(e1) * 1e-3)}
        field.train(was_training)
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--samples", type=int, default=262144)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-codec", action="store_true", help="skip the entropy encode/decode measurement (metric ii)")
    ap.add_argument("--train-steps", type=int, default=5, help="steps of the fwd+bwd(+all-reduce) measurement, 0 = skip")
    ap.add_argument("--no-e2e", action="store_true", help="profiling aid: only the device-resident step is launched (e2e = null)")
    a = ap.parse_args()
    a.warmup = max(a.warmup, 3)

    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product path has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist = None
    if world > 1:
        import torch.distributed as dist

        dist.init_process_group("nccl", device_id=dev)

    Ns = a.samples
    if a.impl == "reference":
        from oracle import ref_py

        if not ref_py.available():
            if rank == 0:
                print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref (reference binaries + bytecode) not built"}))
            return
    arm = Arm(a.impl, dev)
    field = arm.field(seed=0)  # replicas: identical weights on every rank
    pos, dirs = make_inputs(Ns, seed=1000 + rank, device=dev)  # each rank its own sample shard
    model = field
    model.eval()

    from cnc_b200 import _lib

    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    pin_pos, pin_dir = pos.cpu().pin_memory(), dirs.cpu().pin_memory()
    pin_out = torch.empty(Ns, 4, dtype=torch.float32).pin_memory()
    pin_rgb, pin_sig = torch.empty(Ns, 3, dtype=torch.float32).pin_memory(), torch.empty(Ns, 1, dtype=torch.float32).pin_memory()

    def step():
        with torch.no_grad():
            rgb, sigma = model(pos, dirs)
        return rgb, sigma

    def step_e2e():
        if a.impl == "ours":   # the host-buffer entry point of the public API: pinned in, pinned out, copies pipelined
            field.forward_host(pin_pos, pin_dir, pin_rgb, pin_sig)
            return
        with torch.no_grad():
            p = pin_pos.to(dev, non_blocking=True)
            d = pin_dir.to(dev, non_blocking=True)
            rgb, sigma = model(p, d)
            pin_out.copy_(torch.cat([rgb, sigma], -1), non_blocking=True)

    def timed(fn, K, clk=None):
        evs = []
        for _ in range(K):
            flush.zero_()  # L2 flush between timed iterations (outside the event-timed span)
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record()
            fn()
            e.record()
            evs.append((s, e))
        if clk is not None:
            clk.sample_now()   # the queue is still draining: a sample that is certainly under load
        torch.cuda.synchronize()
        return sum(s.elapsed_time(e) for s, e in evs)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(a.warmup):
        step()
        if not a.no_e2e:
            step_e2e()
    barrier()
    l0 = _lib.LAUNCHES
    with Clocks(local) as clk:
        time.sleep(0.02)   # let the sampler thread take its first reading
        ms = timed(step, a.steps, clk)
    launches = _lib.LAUNCHES - l0
    barrier()
    ms_e2e = timed(step_e2e, a.steps) if not a.no_e2e else float("nan")
    barrier()
    # the same host-buffer call as a stream of independent batches: two staging slots on two CUDA streams, so that the first
    # upload / last download of a batch run beside the neighbouring kernels.  Timed as one span over all batches (a flush
    # kernel between them would break the overlap it measures; the 6 MB of inputs per batch are new bytes every time,
    # tables and weights are L2 resident in the flushed measurement as well).
    ms_stream = float("nan")
    if a.impl == "ours" and not a.no_e2e:
        pins = [(pin_pos, pin_dir, pin_rgb, pin_sig),
                (pin_pos.clone().pin_memory(), pin_dir.clone().pin_memory(), torch.empty_like(pin_rgb).pin_memory(), torch.empty_like(pin_sig).pin_memory())]
        side = [torch.cuda.Stream(dev), torch.cuda.Stream(dev)]

        def stream_of_batches(K):
            cur = torch.cuda.current_stream(dev)
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record()
            for st in side:
                st.wait_stream(cur)
            for k in range(K):
                with torch.cuda.stream(side[k & 1]):
                    field.forward_host(*pins[k & 1], slot=k & 1)
            for st in side:
                cur.wait_stream(st)
            e.record()
            torch.cuda.synchronize()
            return s.elapsed_time(e)

        stream_of_batches(4)
        ms_stream = sorted(stream_of_batches(a.steps) for _ in range(3))[1]   # median of three spans of a.steps batches
        barrier()

    # dominant kernel alone, CUDA events on its stream: ours = the fused field kernel (one launch = the
    # whole step); reference = its 3D grid gather kernel_grid<float,3,8> (the top non-library kernel)
    if a.no_e2e:
        ms_k = ms
    elif a.impl == "reference":
        enc = model.mlp_base.encoding_xyz
        xn = ((pos + 1.5) / 3.0).contiguous()
        with torch.no_grad():
            for _ in range(3):
                enc(xn)
            ms_k = timed(lambda: enc(xn), a.steps)
    else:
        for _ in range(3):
            field.fused_forward(pos, dirs)
        ms_k = timed(lambda: field.fused_forward(pos, dirs), a.steps)
    # forward + backward over the same sample batch (metric (i), second half: SURVEY 8d): d(sum rgb + sum sigma)
    # w.r.t. every table and MLP parameter; ours = differentiable path, reference = its K1/K2 kernels under autograd
    fwd_bwd = None
    if a.train_steps > 0 and not a.no_e2e:
        model.train()
        params = [p for p in model.parameters() if p.requires_grad]

        def fb():
            for p in params:
                p.grad = None
            rgb, sigma = model(pos, dirs)
            (rgb.sum() + sigma.sum()).backward()

        for _ in range(3):
            fb()
        ms_fb = timed(fb, max(a.train_steps, 3))
        t_fb = torch.tensor([ms_fb], dtype=torch.float64, device=dev)
        if dist is not None:
            dist.all_reduce(t_fb, op=dist.ReduceOp.MAX)
        ms_fb = float(t_fb.item()) / max(a.train_steps, 3)
        fwd_bwd = {"what": "forward + backward of sum(rgb) + sum(sigma) over the step's sample batch, all table and MLP gradients",
                   "ms_per_step": ms_fb, "samples_per_s": world * Ns / (ms_fb * 1e-3)}
        for p in params:
            p.grad = None
        model.eval()
    # secondary sizes / layouts (N_s = 150 000 as well; --n_features 2 of the training scripts: the fused kernel is
    # specialised for F = 8 / 160 neurons, other layouts run K1 on the sign planes + cuBLAS nn.Linear)
    other = None
    if world == 1 and not a.no_e2e:
        model.eval()
        p15, d15 = pos[:150000].contiguous(), dirs[:150000].contiguous()
        with torch.no_grad():
            for _ in range(3):
                model(p15, d15)
            ms15 = timed(lambda: model(p15, d15), a.steps) / a.steps
        f2 = arm.field(seed=0, F=2).eval()
        with torch.no_grad():
            for _ in range(3):
                f2(pos, dirs)
            ms_f2 = timed(lambda: f2(pos, dirs), a.steps) / a.steps
        other = {"samples_150000": {"ms_per_step": ms15, "samples_per_s": 150000 / (ms15 * 1e-3)},
                 "n_features_2": {"what": "same 12 + 3 x 4 level layout with F = 2 (MLP 87-160-20 / 35-160-160-3), 262 144 samples"
                                          + ("; NOT the fused kernel: GridEncoder on sign planes (K1) + torch nn.Linear" if a.impl == "ours" else ""),
                                  "ms_per_step": ms_f2, "samples_per_s": Ns / (ms_f2 * 1e-3)}}
        del f2
        torch.cuda.empty_cache()
    train = None
    if a.train_steps > 0 and not a.no_e2e and (a.impl == "ours" or world == 1):
        train = train_bench(arm, rank, world, a.train_steps, field)
    t = torch.tensor([ms, ms_e2e, ms_k, ms_stream], dtype=torch.float64, device=dev)
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms, ms_e2e, ms_k, ms_stream = t.tolist()
    codec = None
    if not a.no_codec and not a.no_e2e and (a.impl == "ours" or world == 1):
        del model
        field = None
        torch.cuda.empty_cache()
        codec = codec_bench(arm, rank=rank, world=world, dist=dist)
    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return

    hbm_peak, which = peaks()
    try:
        tf_peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["bf16_tflops"])
    except Exception:
        tf_peak = 1590.0
    s_k = ms_k / a.steps * 1e-3
    if a.impl == "ours":
        ach = Ns * FLOP_PER_SAMPLE_FWD / s_k / 1e12
        roof = {"bound": "tensor", "kernel": "cnc::ff::field_fwd_kernel<false,false> (encode + 5 FC layers, error-compensated tf32 tcgen05)",
                "achieved": ach, "peak": tf_peak, "peak_source": which + " (dense bf16 cuBLAS burst)", "unit": "TFLOP/s",
                "frac": ach / tf_peak, "traffic": ncu_traffic("field_fwd_kernel", Ns),
                "algorithmic_flops_per_launch": Ns * FLOP_PER_SAMPLE_FWD, "ms_per_launch": ms_k / a.steps,
                "note": "algorithmic fp32 FLOPs (189760/sample). fp32 parity costs two MMAs per k-step (kind::tf32 K=8 "
                        "for hi*hi, kind::f16 K=16 on bf16 pairs for the two correction terms), each at half the dense "
                        "bf16 rate, plus K padding (255->256, 95->96); the 160->3 layer runs as FFMA: frac 0.25 is "
                        "the ceiling of this formulation",
                "executed_mma_tflops_bf16_equiv": ach * 4.0 * (256 * 160 + 160 * 80 + 96 * 160 + 160 * 160) / 94880.0,
                "hbm_algorithmic_GBps": Ns * (12 + 4608 + 12 + 16) / s_k / 1e9, "hbm_peak_GBps": hbm_peak}
    else:
        bytes_k = Ns * (12 + 12 * 8 * 32 + 96 * 4)  # 3468 B/point, 3D part of SURVEY 8(d)
        ach = bytes_k / s_k / 1e9
        roof = {"bound": "hbm", "kernel": "kernel_grid<float,3,8> (reference 3D gather)", "achieved": ach,
                "peak": hbm_peak, "peak_source": which, "unit": "GB/s", "frac": ach / hbm_peak, "traffic": None,
                "algorithmic_bytes_per_launch": bytes_k, "ms_per_launch": ms_k / a.steps}
    line = {
        "metric": "ray-samples/sec (encode+MLP)", "value": world * Ns * a.steps / (ms * 1e-3), "unit": "samples/s",
        "n_gpus": world, "steps": a.steps, "warmup": a.warmup, "ms_per_step": ms / a.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "impl": a.impl,
        "config": {"workload": "configs[1]: product field F=8 (3D 12 lvl T=2^19 + 3 planes x 4 lvl T=2^17, MLP "
                               "255-160-80 / 95-160-160-3), forward sigma+rgb over N_s samples per GPU",
                   "samples_per_gpu": Ns, "parallelism": f"ray-sharded x{world}, replicas, no data-path collective",
                   "l2": "flushed between timed iterations (256 MiB write outside the event-timed span)"},
        "e2e": None if a.no_e2e else {"value": world * Ns * a.steps / (ms_e2e * 1e-3), "unit": "samples/s",
                                      "h2d_bytes_per_step": int(pin_pos.numel() * 4 + pin_dir.numel() * 4),
                                      "d2h_bytes_per_step": int(pin_out.numel() * 4),
                                      **({"two_slot_stream": {"value": world * Ns * a.steps / (ms_stream * 1e-3), "unit": "samples/s",
                                                              "what": "same call, independent batches alternating between two staging "
                                                                      "slots / two CUDA streams, one timed span over all batches (median of 3), "
                                                                      "no flush between them"}}
                                         if ms_stream == ms_stream else {})},
        "gpu_launches": int(launches),
        "clocks": clk.summary(),
        "roofline": roof,
    }
    if a.impl == "reference":
        line["cpu_baseline"] = {"value": line["value"], "unit": "samples/s", "cores": 0, "kind": "reference",
                                "sample": "the unmodified reference classes (oracle/_ref/py) on the reference CUDA kernels "
                                          "(oracle/_ref/*.so) + torch fp32 MLP on the GPU: the reference has no CPU "
                                          "implementation of this path; its CPU piece is the entropy coder (codec.coder)"}
    elif world == 1 and not a.no_cpu_baseline:
        line["cpu_baseline"] = cpu_baseline()
    if other is not None:
        line["other_workloads"] = other
    if fwd_bwd is not None:
        line["fwd_bwd"] = fwd_bwd
    if train is not None:
        line["train_step"] = train
        # the step that HAS an exchange, next to the collective-free forward that `value` is (configs[3] of BASELINE.json)
        line["scaling_with_exchange"] = {"metric": "training-step ray-samples/s over all ranks", "value": train.get("samples_per_s"),
                                         "ms_per_step": train.get("ms_per_step"), "scaling": "weak",
                                         "efficiency_vs_replicas_same_run": train.get("efficiency_vs_replicas"),
                                         "comm_bytes_per_step": train.get("comm_bytes_per_step")}
    if codec is not None:
        line["codec"] = codec
    print(json.dumps(line))
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
