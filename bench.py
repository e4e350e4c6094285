#!/usr/bin/env python
"""bench.py — clouds/s of the full VoteNet inference forward (BASELINE.json configs[3]: backbone + vote module +
proposal head + 3-D NMS, batch 8 x 20 000 points (xyz + height) per GPU) on N B200s.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P bench.py --gpus N ...

One "step" = one pass of the hot path over one batch of 8 synthetic SUN-RGB-D-shaped clouds per rank (weak scaling;
at N>1 one NCCL all-gather of the per-rank detection records + the global merge are inside the step).
Prints ONE JSON line (rank 0).  --impl reference times the CPU oracle (the reference's algorithms restated, see
oracle/) on the host cores instead — the only other place besides cpu_baseline where oracle/ is executed.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
# several steps are in flight on independent streams: give every stream its own hardware queue (default is 8)
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")

import numpy as np  # noqa: E402
import torch  # noqa: E402

CLOUDS_PER_RANK = 8
METRIC = "clouds/sec (20k pts, B=8) full VoteNet fwd"
UNIT = "clouds/s"


def _peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm=d["hbm_gbs"], tc_burst=d["bf16_tflops"], tc_sustained=d["bf16_tflops_sustained"], src="measured")
    return dict(hbm=6650.0, tc_burst=1590.0, tc_sustained=1400.0, src="fallback")


class ClockSampler(threading.Thread):
    """SM clock + throttle reasons sampled DURING the timed region (B200_PROFILING.md): NVML every 5 ms (pynvml), with
    `nvidia-smi --query-gpu` polling as the fallback when NVML cannot be loaded."""

    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
    NAMES = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.stop_flag, self.rows, self.max_mhz, self.source = index, threading.Event(), [], None, "nvml"
        self.h = None
        try:
            import pynvml

            pynvml.nvmlInit()
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = index
            if vis:
                try:
                    phys = int(vis.split(",")[index])
                except ValueError:
                    phys = index
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
        except Exception:
            self.h, self.source = None, "nvidia-smi"

    def _sample_nvml(self):
        nv = self.nv
        mhz = float(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
        try:
            r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
        except Exception:
            r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
        bits = [nv.nvmlClocksEventReasonHwSlowdown, nv.nvmlClocksEventReasonHwThermalSlowdown,
                nv.nvmlClocksEventReasonSwThermalSlowdown, nv.nvmlClocksEventReasonSwPowerCap]
        self.rows.append((mhz, [bool(r & b) for b in bits]))

    def _sample_smi(self):
        o = subprocess.run(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits"],
                           capture_output=True, text=True, timeout=5).stdout.strip()
        if o:
            f = [x.strip() for x in o.split(",")]
            self.max_mhz = float(f[1])
            self.rows.append((float(f[0]), [x.lower().startswith("active") for x in f[2:6]]))

    def run(self):
        while not self.stop_flag.is_set():
            try:
                if self.h is not None:
                    self._sample_nvml()
                else:
                    self._sample_smi()
            except Exception:
                pass
            self.stop_flag.wait(0.005 if self.h is not None else 0.2)

    def summary(self):
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": ["unavailable"], "samples": 0, "source": self.source}
        sm = [r[0] for r in self.rows]
        reasons = [n for i, n in enumerate(self.NAMES) if any(r[1][i] for r in self.rows)]
        return {"sm_mhz": float(np.median(sm)), "sm_min_mhz": float(min(sm)), "sm_max_mhz": self.max_mhz, "reasons": reasons,
                "samples": len(self.rows), "source": self.source}


# ----------------------------------------------------------------------------------------------------------------------
def run_reference(args, rank, world):
    """CPU arm: the oracle restatement of the reference's path on the box's host cores, all threads it can use."""
    if rank != 0:
        return
    from concurrent.futures import ThreadPoolExecutor

    from oracle import dense as odense
    from votenet_b200 import synth
    from votenet_b200.config import VoteNetConfig
    from votenet_b200.weights import make_synthetic_weights

    cfg = VoteNetConfig()
    w = make_synthetic_weights(cfg, 0)
    cores = len(os.sched_getaffinity(0))
    torch.set_num_threads(1)
    clouds = [synth.synthetic_cloud(i, cfg.num_points) for i in range(CLOUDS_PER_RANK)]

    def one(i):
        x = clouds[i % len(clouds)][None]
        odense.votenet_forward(x, synth.height_feature(x), w, cfg, synth.CLASS_MEAN_SIZE)

    t0 = time.time(); one(0); t1 = time.time() - t0
    workers = cores
    # A step of this arm is the SAME batch as the GPU arm's: 8 clouds through the full forward, one cloud per host thread
    # (the reference's CPU ops are single-threaded); the clouds of all steps are fed to a pool of every host core, so the
    # cores stay busy across step boundaries.
    per_step = CLOUDS_PER_RANK
    with ThreadPoolExecutor(workers) as ex:
        list(ex.map(one, range(min(args.warmup, 2) * per_step)))
        t0 = time.time()
        list(ex.map(one, range(args.steps * per_step)))
        dt = time.time() - t0
    value = per_step * args.steps / dt
    sample = (f"{args.steps} steps x {CLOUDS_PER_RANK} clouds (full forward incl. NMS, 20000 pts), one cloud per thread on a pool of "
              f"{workers} host threads (every core in sched_getaffinity); single-thread {1.0 / t1:.3f} clouds/s ({t1:.2f} s/cloud)")
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "configs[3]: full VoteNet fwd, 8 clouds x 20000 pts (xyz+height), CPU oracle of the reference path"},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": workers, "kind": "port", "sample": sample,
                             "single_thread_value": 1.0 / t1},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------------------------------------------------
def cpu_baseline(cfg, w, budget_s=20.0):
    """The CPU restatement of the reference path (oracle/), timed on THIS box's host cores: one cloud per thread on
    every core the process may use (the reference's CPU ops are single-threaded, tf_interpolate.cpp / tf_nms3d.cpp), plus
    the single-thread figure."""
    from concurrent.futures import ThreadPoolExecutor

    from oracle import dense as odense
    from votenet_b200 import synth

    cores = len(os.sched_getaffinity(0))
    torch.set_num_threads(1)
    base = [synth.synthetic_cloud(i, cfg.num_points)[None] for i in range(CLOUDS_PER_RANK)]

    def one(i):
        x = base[i % len(base)]
        odense.votenet_forward(x, synth.height_feature(x), w, cfg, synth.CLASS_MEAN_SIZE)

    t0 = time.time(); one(0); t1 = time.time() - t0
    rounds = max(1, min(3, int(budget_s / max(t1, 1e-3) / 2)))
    with ThreadPoolExecutor(cores) as ex:
        t0 = time.time()
        list(ex.map(one, range(cores * rounds)))
        dt = time.time() - t0
    torch.set_num_threads(cores)
    return {"value": cores * rounds / dt, "unit": UNIT, "cores": cores, "kind": "port",
            "single_thread_value": 1.0 / t1,
            "sample": f"{cores * rounds} clouds of the same workload ({rounds} rounds x {cores} host threads = every core in "
                      f"sched_getaffinity, 1 cloud per thread); single-thread {1.0 / t1:.3f} clouds/s ({t1:.2f} s/cloud)"}


def kernel_table(eng, cfg, peaks, flush):
    """Per-stage device time (CUDA events on the launching stream, L2 flushed before each launch, each stage launched
    alone) and roofline fraction from ALGORITHMIC bytes / flops (SURVEY.md §8(d) per-cloud figures x batch)."""
    from votenet_b200._lib import check, dptr, lib, stream_ptr

    B = eng.B
    s = eng.slots[0]
    st = torch.cuda.current_stream()
    sp = stream_ptr

    def t(fn, it=5):
        fn(); torch.cuda.synchronize()
        ts = []
        for _ in range(it):
            flush.zero_()
            a, b = torch.cuda.Event(True), torch.cuda.Event(True)
            a.record(st); fn(); b.record(st); torch.cuda.synchronize()
            ts.append(a.elapsed_time(b))
        return float(np.mean(ts))

    rows = []

    def hbm(name, ms, byt, note=None):
        r = dict(kernel=name, ms=ms, bound="hbm", algo_bytes=byt, achieved=byt / ms / 1e6, peak=peaks["hbm"], unit="GB/s")
        if note:
            r["note"] = note
        rows.append(r)

    def tensor(name, ms, fl, executed=None):
        r = dict(kernel=name, ms=ms, bound="tensor", algo_flops=fl, achieved=fl / ms / 1e9, peak=peaks["tc_burst"], unit="TFLOP/s")
        if executed is not None:   # MMA flops the kernels really issue (tile packing drops the padded duplicate rows, layer 1
            r["executed_flops"] = executed            # of the wide levels runs once per source point instead of per grouped row)
            r["frac_executed"] = executed / ms / 1e9 / peaks["tc_burst"]
        rows.append(r)

    def sa_executed(cnt, n_src, cin_feat, widths, hoisted):
        """Executed tensor-core flops of one fused SA stage: every centroid occupies a slot of 16 / 32 / 64 rows."""
        c = cnt.clamp(min=1)
        slot_rows = torch.where(c <= 16, 16, torch.where(c <= 32, 32, 64)).sum().item()
        c1, c2, c3 = widths
        if hoisted:   # layer 1 = one (B*n_src, cin_feat) x (cin_feat, c1) GEMM + a rank-3 fp32 update in the producer
            return 2.0 * (B * n_src * cin_feat * c1 + slot_rows * (c1 * c2 + c2 * c3))
        return 2.0 * slot_rows * (16 * c1 + c1 * c2 + c2 * c3)   # K of layer 1 padded to 16

    def mlp_flops(rows_, cin, widths):
        fl = 0
        for co in widths:
            fl += cin * co; cin = co
        return 2.0 * rows_ * fl

    onchip = "on-chip kernel: achieved = streaming-equivalent bytes (m-1)*n*16 B per cloud / time (may exceed the HBM peak)"
    src, c, feat = s.xyz, cfg.feature_dim, s.feat
    for li, l in enumerate(s.lv):
        sa = cfg.sa[li]
        if li == 0:
            ms = t(lambda: check(lib.vnb_farthest_point_sample(B, l.n, l.m, dptr(src), dptr(l.fps), sp())))
            hbm("fps_sa1", ms, B * (l.m - 1) * l.n * 16, onchip)
        elif li == 1:  # FPS of an FPS-ordered set: parallel proof of the identity prefix (sequential kernel only on failure)
            ms = t(lambda: check(lib.vnb_farthest_point_sample_nested_proof(B, l.n, l.m, dptr(src), dptr(l.fps), dptr(s.fps_ws),
                                                                            dptr(s.fps_tie), sp())))
            rows.append(dict(kernel="fps_sa2_identity_proof", ms=ms, bound="hbm", algo_bytes=B * l.n * 12, achieved=B * l.n * 12 / ms / 1e6,
                             peak=peaks["hbm"], unit="GB/s",
                             note="proof kernels: the sampled set is a prefix of sa1's picks; latency-bound (compulsory bytes only)"))
        else:  # covered by the sa2-level proof: identity prefix written by early-exit launches (no sampling work at all)
            ms = t(lambda: check(lib.vnb_farthest_point_sample_nested_hint(B, l.n, l.m, dptr(src), dptr(l.fps), dptr(s.fps_ws),
                                                                           dptr(s.fps_tie), sp())))
            rows.append(dict(kernel=f"fps_sa{li + 1}_covered_by_proof", ms=ms, bound="hbm", algo_bytes=B * l.m * 4,
                             achieved=B * l.m * 4 / ms / 1e6, peak=peaks["hbm"], unit="GB/s",
                             note="no sampler runs (identity prefix proven at sa2): launch latency of the early-exit kernels"))
        ms = t(lambda: check(lib.vnb_query_ball_point_ws(B, l.n, l.m, float(sa.radius), 64, dptr(src), dptr(l.xyz), dptr(l.idx),
                                                         dptr(l.cnt), dptr(l.bq_ws), sp())))
        hbm(f"ball_query_sa{li + 1}", ms, B * (l.n * 12 + l.m * 12 + l.m * 256 + l.m * 4), "latency-bound (compulsory bytes only)")
        if l.bq_ws is not None and l.n >= 1024:   # grid path: the two halves of the call (the engine builds sa1's grid beside the FPS)
            ms_b = t(lambda: check(lib.vnb_query_ball_point_prepare(B, l.n, float(sa.radius), dptr(src), dptr(l.bq_ws), sp())))
            ms_q = t(lambda: check(lib.vnb_query_ball_point_prepared(B, l.n, l.m, float(sa.radius), 64, dptr(src), dptr(l.xyz),
                                                                     dptr(l.idx), dptr(l.cnt), dptr(l.bq_ws), sp())))
            rows[-1].update(ms_grid_build=ms_b, ms_query=ms_q)
            rows[-1]["note"] += "; ms = grid build (one CTA per cloud; needs the searched set only) + query"
        ms = t(lambda: eng._sa(li, src, feat, l.n, c, l.xyz, l.idx, l.m, l.q, l.feat, st, s.sa_ws, l.cnt))
        tensor(f"sa{li + 1}_group_mlp_max", ms, mlp_flops(B * l.m * 64, 3 + c, sa.mlp),
               sa_executed(l.cnt, l.n, c, sa.mlp, eng.sa_layers[li][3] is not None))
        src, feat, c = l.xyz, l.feat, sa.mlp[-1]
    # feature propagation: three_nn + (interpolate/concat + 2 x (1x1 conv + BN + ReLU)); voting module
    from votenet_b200.utils import fp_module_fused
    ns, cf = s.lv[1].m, cfg.seed_feat_dim
    for f, scope, u, kx in zip(s.fp, ("fp1", "fp2"), (s.lv[2].xyz, s.lv[1].xyz), (s.lv[3].xyz, s.lv[2].xyz)):
        ms = t(lambda: check(lib.vnb_three_nn(B, f.n, f.m, dptr(u), dptr(kx), dptr(f.dist), dptr(f.idx), sp())))
        hbm(f"three_nn_{scope}", ms, B * (f.n * 12 + f.m * 12 + f.n * 24), "latency-bound (compulsory bytes only)")
    f1, f2 = s.fp
    fl_fp1 = mlp_flops(B * f1.n, f1.c1 + f1.c2, cfg.fp_mlp)
    fl_fp2 = mlp_flops(B * f2.n, f2.c1 + f2.c2, cfg.fp_mlp)
    fl_vote = mlp_flops(B * ns, 3 + cf, cfg.vote_units)
    if eng.fuse_fp:
        ms = t(lambda: fp_module_fused(f1.dist, f1.idx, s.lv[2].feat, s.lv[3].feat,
                                       [eng.store.layer(f"fp1/conv_{i}") for i in range(2)], f1.h[-1], stream=st))
        tensor("fp1_interpolate_mlp_fused", ms, fl_fp1)
        ms = t(lambda: fp_module_fused(f2.dist, f2.idx, s.lv[1].feat, f1.h[-1].view(B, f1.n, -1),
                                       [eng.store.layer(f"fp2/conv_{i}") for i in range(2)], None,   # as the engine calls it
                                       vote=(eng.vote_fused, eng.vote_x0, s.lv[1].xyz, s.votes_xyz, s.votes_feat), stream=st))
        tensor("fp2_interpolate_mlp_vote_fused", ms, fl_fp2 + fl_vote)
    else:
        pts2 = s.lv[3].feat
        for f, skip, scope, fl in zip(s.fp, (s.lv[2].feat, s.lv[1].feat), ("fp1", "fp2"), (fl_fp1, fl_fp2)):
            def fp_fn(f=f, skip=skip, scope=scope, pts2=pts2):
                check(lib.vnb_fp_interpolate_concat(B, f.n, f.m, f.c1, f.c2, dptr(f.dist), dptr(f.idx), dptr(skip), dptr(pts2),
                                                    dptr(f.cat), sp()))
                x = f.cat
                for i in range(len(cfg.fp_mlp)):
                    eng._linear(B * f.n, x, eng.store.layer(f"{scope}/conv_{i}"), True, f.h[i], None, st)
                    x = f.h[i]
            ms = t(fp_fn)
            tensor(f"{scope}_interpolate_mlp", ms, fl)
            pts2 = f.h[-1]

        def vote_fn():
            check(lib.vnb_concat2(B * ns, 3, cf, dptr(s.lv[1].xyz), dptr(pts2), dptr(s.seeds), sp()))
            x = s.seeds
            nv = len(cfg.vote_units)
            for i in range(nv):
                eng._linear(B * ns, x, eng.store.layer(f"voting{i}"), i < nv - 1, s.vh[i], None, st,
                            residual=s.seeds if i == nv - 1 else None)
                x = s.vh[i]
            check(lib.vnb_split2(B * ns, 3, cf, dptr(x), dptr(s.votes_xyz), dptr(s.votes_feat), sp()))
        ms = t(vote_fn)
        tensor("vote_mlp", ms, fl_vote)
    p = cfg.proposal

    def prop_fn():
        check(lib.vnb_farthest_point_sample_nested_hint(B, ns, p.npoint, dptr(s.lv[1].xyz), dptr(s.p_fps), dptr(s.fps_ws),
                                                        dptr(s.fps_tie), sp()))
        check(lib.vnb_gather_point(B, ns, p.npoint, dptr(s.votes_xyz), dptr(s.p_fps), dptr(s.p_xyz), sp()))
        check(lib.vnb_query_ball_point(B, ns, p.npoint, float(p.radius), 64, dptr(s.votes_xyz), dptr(s.p_xyz), dptr(s.p_idx),
                                       dptr(s.p_cnt), sp()))
        eng._sa(len(cfg.sa), s.votes_xyz, s.votes_feat, ns, cf, s.p_xyz, s.p_idx, p.npoint, s.p_q, s.p_feat, st, s.sa_ws, s.p_cnt)
        x = s.p_feat
        for i in range(len(p.mlp2)):
            eng._linear(B * p.npoint, x, eng.store.layer(f"proposal/conv_post_{i}"), i < len(p.mlp2) - 1, s.p_h[i], None, st)
            x = s.p_h[i]
    ms = t(prop_fn)
    tensor("proposal_fps_ballquery_group_mlp", ms, mlp_flops(B * p.npoint * 64, 3 + cf, p.mlp) + mlp_flops(B * p.npoint, p.mlp[-1], p.mlp2))
    r = s.rec

    def nms_fn():
        check(lib.vnb_decode_nms3d(B, p.npoint, dptr(s.p_xyz), dptr(s.p_h[-1]), dptr(eng.mean_size), float(cfg.nms_iou),
                                   dptr(r.bboxes), dptr(r.scores), dptr(r.objectness), dptr(r.class_scores), dptr(r.keep),
                                   dptr(r.nms_idx), dptr(r.nms_key), dptr(r.nms_count), dptr(s.bboxes_pred),
                                   dptr(s.class_scores_pred), dptr(s.batch_idx), dptr(s.nms_ws), sp()))
    ms = t(nms_fn)
    hbm("decode_nms3d", ms, B * (p.npoint * (79 + 3) * 4 + eng.record_nbytes // B), "latency-bound (compulsory bytes only)")
    for r_ in rows:
        r_["frac"] = r_["achieved"] / r_["peak"]
    return rows


def latency_inflight1(eng, xyz, feat, dev, reps=12):
    """One forward alone on an idle GPU (graph replay, synchronised on both sides): median latency in ms."""
    st = torch.cuda.Stream(device=dev)
    ts = []
    for _ in range(reps):
        torch.cuda.synchronize(dev)
        a, b = torch.cuda.Event(True), torch.cuda.Event(True)
        a.record(st)
        eng.infer_device(xyz, feat, stream=st)
        b.record(st)
        torch.cuda.synchronize(dev)
        ts.append(a.elapsed_time(b))
    return float(np.median(ts[2:]))


def configs1_single_sa(eng, cfg, dev, flush, reps=5):
    """BASELINE.json configs[1]: ONE set-abstraction layer on the raw clouds — FPS 20000 -> 1024, ball query r = 0.2,
    nsample 64, group + shared MLP [64, 64, 128] + max-pool — the four launches back to back, L2 flushed before each run."""
    from votenet_b200._lib import check, dptr, lib, stream_ptr

    B, N, m = eng.B, eng.N, 1024
    s = eng.slots[0]
    i32 = torch.int32
    fps = torch.empty((B, m), dtype=i32, device=dev)
    nxyz = torch.empty((B, m, 3), device=dev)
    idx = torch.zeros((B, m, 64), dtype=i32, device=dev)
    cnt = torch.empty((B, m), dtype=i32, device=dev)
    out = torch.empty((B, m, cfg.sa[0].mlp[-1]), device=dev)
    st = torch.cuda.current_stream(dev)
    ev = {}

    def run(timed):
        marks = [torch.cuda.Event(True) for _ in range(5)]
        marks[0].record(st)
        check(lib.vnb_farthest_point_sample(B, N, m, dptr(s.xyz), dptr(fps), stream_ptr()))
        marks[1].record(st)
        check(lib.vnb_gather_point(B, N, m, dptr(s.xyz), dptr(fps), dptr(nxyz), stream_ptr()))
        check(lib.vnb_query_ball_point_ws(B, N, m, 0.2, 64, dptr(s.xyz), dptr(nxyz), dptr(idx), dptr(cnt), dptr(s.lv[0].bq_ws),
                                          stream_ptr()))
        marks[2].record(st)
        eng._sa(0, s.xyz, s.feat, N, cfg.feature_dim, nxyz, idx, m, None, out, st, s.sa_ws, cnt)
        marks[3].record(st)
        torch.cuda.synchronize(dev)
        if timed:
            for k, (i, j) in dict(total=(0, 3), fps=(0, 1), ball_query=(1, 2), group_mlp_max=(2, 3)).items():
                ev.setdefault(k, []).append(marks[i].elapsed_time(marks[j]))

    run(False)
    for _ in range(reps):
        flush.zero_()
        run(True)
    ms = {k: float(np.median(v)) for k, v in ev.items()}
    return {"workload": "configs[1]: single SA layer, FPS 20000->1024, ball r=0.2 nsample=64, group + MLP [64,64,128] + max, "
                        f"{B} clouds (xyz + height)", "ms": ms["total"], "clouds_per_s": B / (ms["total"] / 1e3),
            "ms_fps": ms["fps"], "ms_ball_query": ms["ball_query"], "ms_group_mlp_max": ms["group_mlp_max"]}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=600)
    ap.add_argument("--warmup", type=int, default=24)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--precision", type=int, default=1)
    ap.add_argument("--inflight", type=int, default=16, help="steps in flight (independent workspaces + streams)")
    ap.add_argument("--fps-cluster", type=int, default=4,
                    help="CTAs per FPS cluster (4 = 32 SMs per batch: best throughput with steps overlapped; 8 = lowest latency)")
    ap.add_argument("--fps-variant", type=int, default=None, help="0: register/cluster FPS kernel, 1: bucket-pruned (default)")
    ap.add_argument("--fps-threads", type=int, default=None, help="tuning: threads per FPS CTA (256/512/1024)")
    ap.add_argument("--tune", action="append", default=[], metavar="KEY=VALUE", help="vnb_set_tuning knob (repeatable)")
    ap.add_argument("--transport", default="peer", choices=["peer", "nccl"],
                    help="N>1 exchange of the detection records: one-sided pushes over NVLink peer memory, or an NCCL all-gather")
    ap.add_argument("--debug-skip-fps1", action="store_true", help="experiment only: reuse the warm-up step's SA1 FPS result")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    assert torch.cuda.is_available(), "bench.py needs a GPU (no CPU fallback); use --impl reference for the CPU arm"
    assert world == args.gpus or world == 1 and args.gpus == 1, f"--gpus {args.gpus} but WORLD_SIZE={world}"
    import torch.distributed as dist

    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    from votenet_b200 import synth
    from votenet_b200._lib import lib
    from votenet_b200.config import VoteNetConfig
    from votenet_b200.dist import make_gather, shard_range
    from votenet_b200.engine import Engine
    from votenet_b200.weights import make_synthetic_weights

    from votenet_b200._lib import check as _check
    if args.fps_cluster is not None:
        _check(lib.vnb_set_tuning(b"fps_cluster", args.fps_cluster))
    if args.fps_threads is not None:
        _check(lib.vnb_set_tuning(b"fps_threads", args.fps_threads))
    if args.fps_variant is not None:
        _check(lib.vnb_set_tuning(b"fps_variant", args.fps_variant))
    fps_variant = 1 if args.fps_variant is None else args.fps_variant
    trap_buf = torch.zeros(8, dtype=torch.int32).pin_memory()  # filled by a device-side bounded wait that gives up
    _check(lib.vnb_debug_trap_buffer(trap_buf.data_ptr()))
    global TRAP_BUF
    TRAP_BUF = trap_buf
    for kv in args.tune:
        k, v = kv.split("=")
        _check(lib.vnb_set_tuning(k.encode(), int(v)))
    peaks = _peaks()
    cfg = VoteNetConfig()  # BASELINE.json: 20 000 points, (xyz + height)
    B, N = CLOUDS_PER_RANK, cfg.num_points
    w = make_synthetic_weights(cfg, 0)
    eng = Engine(cfg, w, B, device=dev, precision=args.precision, use_graph=not args.no_graph, slots=args.inflight)
    eng.debug_skip_fps1 = args.debug_skip_fps1

    # ---- synthetic inputs: this rank's clouds; a ring of RING device-resident batches (164 MB > the 126 MB L2) so that
    #      a step's inputs are never L2-resident from an earlier step
    ids = list(shard_range(rank, world, B))
    base = np.stack([synth.synthetic_cloud(i, N) for i in ids], 0)
    RING = 64
    ring_xyz, ring_feat = [], []
    bx = torch.as_tensor(base, device=dev)
    for r in range(RING):
        a = 2 * np.pi * r / RING
        rot = torch.tensor([[np.cos(a), 0, np.sin(a)], [0, 1, 0], [-np.sin(a), 0, np.cos(a)]], dtype=torch.float32, device=dev)
        x = (bx @ rot.T).contiguous() if r else bx.clone()
        ring_xyz.append(x)
        ring_feat.append((synth.FLOOR_Y - x[..., 1:2]).contiguous())
    HR = max(4, args.inflight)
    host_xyz = [ring_xyz[r].cpu().pin_memory() for r in range(HR)]
    host_feat = [ring_feat[r].cpu().pin_memory() for r in range(HR)]
    host_out = [torch.empty((eng.record_nbytes,), dtype=torch.uint8).pin_memory() for _ in range(HR)]
    h2d = host_xyz[0].numel() * 4 + host_feat[0].numel() * 4
    d2h = eng.record_nbytes + (world * B * cfg.proposal.npoint * 8 + 4 if world > 1 else 0)

    NS = args.inflight
    gather, transport = (make_gather(world, rank, B, cfg.proposal.npoint, dev, slots=NS, transport=args.transport)
                         if world > 1 else (None, "none"))
    host_merged = [torch.empty((world * B * cfg.proposal.npoint * 2 + 1,), dtype=torch.int32).pin_memory() for _ in range(HR)] if world > 1 else None
    streams = [torch.cuda.Stream(device=dev) for _ in range(NS)]
    ctl = torch.cuda.current_stream(dev)

    def step_device(i):
        st = streams[eng._step % NS]  # a slot always runs on the same stream (slot = eng._step % inflight)
        rec = eng.infer_device(ring_xyz[i % RING], ring_feat[i % RING], stream=st)
        if world > 1:
            with torch.cuda.stream(st):
                gather(rec.buf, slot=(eng._step - 1) % NS)

    def step_host(i):
        st = streams[eng._step % NS]
        eng.infer_host(host_xyz[i % HR], host_feat[i % HR], host_out[i % HR], stream=st)
        if world > 1:   # the step's result is the merged (Nnms,2) list of the WHOLE batch + its length: read it back too
            with torch.cuda.stream(st):
                idx, cnt = gather(eng.slots[(eng._step - 1) % NS].rec.buf, slot=(eng._step - 1) % NS)[:2]
                hm = host_merged[i % HR]
                hm[:-1].copy_(idx.view(-1), non_blocking=True)
                hm[-1:].copy_(cnt, non_blocking=True)

    def timed(step_fn, K, W):
        for i in range(W):
            step_fn(i)
        for st in streams:
            ctl.wait_stream(st)
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)
        n0 = lib.vnb_launch_count()
        e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
        t0 = time.time()
        e0.record(ctl)
        for st in streams:
            st.wait_event(e0)
        for i in range(K):
            step_fn(W + i)
        enq = time.time() - t0
        for st in streams:
            ctl.wait_stream(st)
        e1.record(ctl)
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)
        wall = time.time() - t0
        ms = e0.elapsed_time(e1)
        if world > 1:
            tt = torch.tensor([ms], device=dev)
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            ms = float(tt.item())
        return ms, wall, enq

    # one-time setup, not a warm-up step: every slot captures its CUDA graph on first use (an eager pass + the capture,
    # tens of ms), which must not land inside the timed region when --warmup is smaller than --inflight
    for i in range(len(eng.slots)):
        step_device(i)
    for st in streams:
        ctl.wait_stream(st)
    torch.cuda.synchronize(dev)
    eng._step = 0

    sampler = ClockSampler(local)
    sampler.start()
    ms_dev, wall_dev, enq_dev = timed(step_device, args.steps, args.warmup)
    sampler.stop_flag.set()
    ms_e2e, wall_e2e, enq_e2e = timed(step_host, args.steps, args.warmup)
    lpf = eng.launches_per_forward or 0
    per_step_extra = 0 if world == 1 else (3 if transport == "peer" else 1)   # push + wait + merge | merge (NCCL's kernel is not ours)
    launches = (lpf + per_step_extra) * args.steps

    value = world * B * args.steps / (ms_dev / 1e3)
    e2e = world * B * args.steps / (ms_e2e / 1e3)

    if rank == 0:
        flush = torch.empty(256 * 1024 * 1024 // 4, device=dev)
        ktab = kernel_table(eng, cfg, peaks, flush)
        dom = max(ktab, key=lambda r: r["ms"])
        traffic = None
        tp = os.path.join(ROOT, "profiles", "roofline_traffic.json")
        if os.path.exists(tp):
            traffic = json.load(open(tp)).get(dom["kernel"])
        roof = {"kernel": dom["kernel"], "bound": "hbm" if dom["unit"] == "GB/s" else "tensor", "achieved": dom["achieved"],
                "peak": dom["peak"], "unit": dom["unit"], "frac": dom["frac"], "traffic": traffic,
                "peak_source": peaks["src"], "ms_per_launch": dom["ms"],
                # share of one forward's SERIALISED kernel time (what the ncu launch list in profiles/ shows); forwards
                # overlap in the timed loop, so ms_per_launch is larger than ms_per_step
                "share_of_serialised_forward": dom["ms"] / sum(r["ms"] for r in ktab),
                "note": dom.get("note", "")}
        lat1 = latency_inflight1(eng, ring_xyz[1], ring_feat[1], dev)
        cfg1 = configs1_single_sa(eng, cfg, dev, flush)
        cpu = None if args.no_cpu_baseline else cpu_baseline(cfg, w)
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": ms_dev / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "f16 operands / f32 accumulate (tcgen05); f32 geometry; int32 indices" if args.precision == 1 else "f32",
                "data": "synthetic",
                "config": {"workload": "configs[3]: full VoteNet fwd (4 SA + 2 FP + vote + proposal + decode + 3D-NMS), "
                                       f"{B} clouds x {N} pts (xyz+height) per GPU", "clouds_per_gpu": B, "points": N,
                           "parallelism": f"dp{world} (clouds sharded, no data-path collective; per forward ONE all-gather of the detection records, transport: {transport})" if world > 1 else "single GPU",
                           "l2": f"inputs rotate over {RING} device-resident batches ({RING * h2d / 1e6:.0f} MB > 126 MB L2)",
                           "pipelining": f"{NS} steps in flight ({NS} workspaces, {NS} streams): later steps' FPS chains overlap earlier steps' MLP chains",
                           "cuda_graph": not args.no_graph,
                           "fps_geometry": ("bucket-pruned sampler, ONE CTA of 512 threads per cloud (8 SMs per batch); nested levels: one "
                                            "parallel identity-prefix proof at sa2, deeper levels covered by it") if fps_variant == 1
                                           else f"register/cluster sampler, {args.fps_threads or 256} threads x cluster {args.fps_cluster}",
                           "float_tolerance": "1e-3 per module on identical inputs, measured as max|a-b| / max|b| (relative to the tensor's scale, "
                                              "not element-wise); observed per-stage errors: profiles/r2_stage_errors.json (written by tests/test_gpu_forward.py)"},
                "e2e": {"value": e2e, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                        "ms_per_step": ms_e2e / args.steps},
                "gpu_launches": launches, "launches_per_forward": lpf,
                # one forward alone on the idle GPU: the throughput above needs several forwards in flight (one forward's
                # sa1 sampler keeps 8 of 148 SMs busy for most of this latency)
                "latency_ms_inflight1": lat1, "clouds_per_s_inflight1": B / (lat1 / 1e3),
                "configs1_single_sa_layer": cfg1,
                "roofline": roof, "kernels": [{k: (round(v, 4) if isinstance(v, float) else v) for k, v in r.items()} for r in ktab],
                "cpu_baseline": cpu, "clocks": sampler.summary(),
                "wall_s": {"device_loop": wall_dev, "e2e_loop": wall_e2e},
                "host_enqueue_ms_per_step": {"device_loop": 1e3 * enq_dev / args.steps, "e2e_loop": 1e3 * enq_e2e / args.steps}}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


TRAP_BUF = None

if __name__ == "__main__":
    try:
        main()
    except Exception:
        if TRAP_BUF is not None:
            print("device-side trap record {line, blockDim, blockIdx, threadIdx, gridDim}:", TRAP_BUF[:5].tolist(),
                  file=sys.stderr, flush=True)
        raise
