"""Pre-allocated, multi-stream, CUDA-graph engine for the VoteNet inference tower (bench.py's hot loop).

Same graph as `model.VoteNetB200.forward` (reference: /root/reference/model.py:34-61,85-137), but
  * every buffer is allocated once for a fixed (batch, num_points); the C ABI is called with raw pointers;
  * the forward is scheduled on three streams — the FPS chain depends on xyz only, ball queries
    and three_nn depend on xyz + centroids only, so both run beside the feature (MLP) chain and are joined by events;
  * the whole multi-stream forward is captured into one CUDA graph per slot;
  * `slots` independent workspaces (bench.py: 16) keep several forwards in flight: the sa1 FPS of a forward occupies
    8 SMs for 1.5 ms, the other 140 SMs run the feature chains of earlier forwards;
  * the nested sampling levels sample prefixes of sa1's picks: the first one proves the identity prefix in parallel
    (no tie tracking in the big sampler), the deeper ones and the proposal module are covered by that proof, the FP
    modules and the voting module run as one fused tensor-core kernel per FP level, the SA kernels get the ball
    query's hit counts and skip the padded duplicate rows;
  * detections are written straight into one contiguous record (the all-gather wire layout).
"""
import ctypes as C

import torch

from ._lib import check, dptr, lib
from .config import NC, PROPOSAL_CHANNELS, VoteNetConfig
from .synth import CLASS_MEAN_SIZE
from .utils import (HOIST_MIN_C, PRECISION_TENSOR, Layer, WeightStore, fp_module_fused, vote_layers_fused)


def _sp(stream):
    return C.c_void_p(stream.cuda_stream)


class DetectionRecord:
    """Fixed-size per-rank detection record: one contiguous byte buffer + typed views (no pack step)."""

    FIELDS = (("bboxes", torch.float32, lambda b, k: (b, k, 8, 3)), ("scores", torch.float32, lambda b, k: (b, k)),
              ("class_scores", torch.float32, lambda b, k: (b, k, NC)), ("objectness", torch.float32, lambda b, k: (b, k, 2)),
              ("keep", torch.uint8, lambda b, k: (b, k)), ("nms_idx", torch.int32, lambda b, k: (b * k, 2)),
              ("nms_key", torch.int32, lambda b, k: (b * k,)), ("nms_count", torch.int32, lambda b, k: (1,)))
    # nms_idx / nms_key / nms_count = this rank's detections as a list sorted by descending score (rows (batch, box)
    # + order-preserving score keys): what vnb_merge_detections merges across ranks without re-sorting

    def __init__(self, b, k, device=None, buf=None):
        self.b, self.k = b, k
        self.offsets = {}
        off = 0
        for name, dt, shp in self.FIELDS:
            n = 1
            for s in shp(b, k):
                n *= s
            nbytes = n * torch.empty((), dtype=dt).element_size()
            self.offsets[name] = (off, nbytes)
            off += (nbytes + 255) // 256 * 256
        self.nbytes = off
        self.buf = buf if buf is not None else torch.zeros((self.nbytes,), dtype=torch.uint8, device=device)
        for name, dt, shp in self.FIELDS:
            o, nb = self.offsets[name]
            setattr(self, name, self.buf[o:o + nb].view(dt).view(shp(b, k)))


class _Slot:
    pass


class Engine:
    def __init__(self, cfg: VoteNetConfig, weights, batch, device="cuda", precision=PRECISION_TENSOR, use_graph=True,
                 slots=2):
        self.cfg, self.B, self.N = cfg, int(batch), cfg.num_points
        self.device = torch.device(device)
        self.precision = precision
        self.use_graph = use_graph
        self.store = WeightStore(weights, device=self.device, eps=cfg.bn_eps, precision=precision)
        self.mean_size = torch.as_tensor(CLASS_MEAN_SIZE, device=self.device).contiguous()
        with torch.cuda.device(self.device):
            self._prepare_layers()
            self.slots = [self._make_slot() for _ in range(slots)]
        self._step = 0
        self.launches_per_forward = None
        # rounds of the sa1 FPS whose arg-max uniqueness is tracked: the largest nested sample count that relies on it
        self.tie_rounds = max([sa.npoint for sa in cfg.sa[1:]] + [cfg.proposal.npoint])
        self.debug_skip_fps1 = False  # experiment only: capture the graph without the SA1 FPS launch
        self.timeline = None   # debugging: set to [] (with use_graph=False) to collect (step, stage, event) marks

    # ------------------------------------------------------------------------------------------------ weights
    def _prepare_layers(self):
        cfg, st = self.cfg, self.store
        self.sa_layers = []
        cfeat = cfg.feature_dim
        for li, sa in enumerate(list(cfg.sa) + [cfg.proposal]):
            scope = f"sa{li + 1}" if li < len(cfg.sa) else "proposal"
            l1, l2, l3 = (st.layer(f"{scope}/conv{i}") for i in range(3))
            c_in = cfeat if li < len(cfg.sa) else cfg.seed_feat_dim
            hoist = self.precision == PRECISION_TENSOR and c_in >= HOIST_MIN_C
            lf = Layer(l1.W[3:].contiguous(), l1.b, self.device) if hoist else None
            if self.precision == PRECISION_TENSOR:
                for l in (l1, l2, l3):
                    l.img  # pack now
                if lf is not None:
                    lf.img
            self.sa_layers.append((l1, l2, l3, lf, c_in))
            cfeat = sa.mlp[-1]
        names = [f"fp1/conv_{i}" for i in range(len(cfg.fp_mlp))] + [f"fp2/conv_{i}" for i in range(len(cfg.fp_mlp))]
        names += [f"voting{i}" for i in range(len(cfg.vote_units))]
        names += [f"proposal/conv_post_{i}" for i in range(len(cfg.proposal.mlp2))]
        for nme in names:
            l = st.layer(nme)
            if self.precision == PRECISION_TENSOR:
                l.img
        # fp modules (+ the voting stack behind the last one) as ONE kernel each when the widths are the fused kernel's
        self.fuse_fp = (self.precision == PRECISION_TENSOR and tuple(cfg.fp_mlp) == (256, 256)
                        and all(sa.mlp[-1] == 256 for sa in cfg.sa[1:]) and tuple(cfg.vote_units) == (256, 256, 259))
        if self.fuse_fp:
            self.vote_fused, self.vote_x0 = vote_layers_fused(st, [f"voting{i}" for i in range(3)])
            for l in self.vote_fused:
                l.img
        torch.cuda.synchronize(self.device)

    # ------------------------------------------------------------------------------------------------ buffers
    def _make_slot(self):
        cfg, B, N, dev = self.cfg, self.B, self.N, self.device
        s = _Slot()
        f32, i32, f16 = torch.float32, torch.int32, torch.float16
        E = lambda shape, dt=f32: torch.empty(shape, dtype=dt, device=dev)  # noqa: E731
        s.xyz = E((B, N, 3))
        s.feat = E((B, N, cfg.feature_dim))
        s.lv = []
        n = N
        for li, sa in enumerate(cfg.sa):
            l = _Slot()
            l.n, l.m = n, sa.npoint
            l.fps = E((B, sa.npoint), i32)
            l.xyz = E((B, sa.npoint, 3))
            l.idx = torch.zeros((B, sa.npoint, sa.nsample), dtype=i32, device=dev)
            l.cnt = E((B, sa.npoint), i32)
            l.q = E((B * n, sa.mlp[0]), f16) if self.sa_layers[li][3] is not None else None
            l.feat = E((B, sa.npoint, sa.mlp[-1]))
            l.bq_ws = (torch.empty((lib.vnb_query_ball_point_workspace_bytes(B, n),), dtype=torch.uint8, device=dev)
                       if n >= 512 else None)   # the C side picks scan / grid by n (tuning: bq_grid_min_n)
            s.lv.append(l)
            n = sa.npoint
        c = cfg.sa[-1].mlp[-1]
        s.fp = []
        for (nu, mk, cskip) in ((cfg.sa[2].npoint, cfg.sa[3].npoint, cfg.sa[2].mlp[-1]),
                                (cfg.sa[1].npoint, cfg.sa[2].npoint, cfg.sa[1].mlp[-1])):
            f = _Slot()
            f.n, f.m, f.c1, f.c2 = nu, mk, cskip, c
            f.dist, f.idx = E((B, nu, 3)), E((B, nu, 3), i32)
            f.cat = E((B * nu, c + cskip))
            f.h = [E((B * nu, co)) for co in cfg.fp_mlp]
            s.fp.append(f)
            c = cfg.fp_mlp[-1]
        ns = cfg.sa[1].npoint
        s.seeds = E((B * ns, 3 + c))
        s.vh = [E((B * ns, co)) for co in cfg.vote_units]
        s.votes_xyz, s.votes_feat = E((B, ns, 3)), E((B, ns, c))
        p = cfg.proposal
        s.p_fps = E((B, p.npoint), i32)
        s.p_xyz = E((B, p.npoint, 3))
        s.p_idx = torch.zeros((B, p.npoint, p.nsample), dtype=i32, device=dev)
        s.p_cnt = E((B, p.npoint), i32)
        s.p_q = E((B * ns, p.mlp[0]), f16) if self.sa_layers[-1][3] is not None else None
        s.p_feat = E((B, p.npoint, p.mlp[-1]))
        s.p_h = [E((B * p.npoint, co)) for co in p.mlp2]
        s.rec = DetectionRecord(B, p.npoint, dev)
        s.nms_ws = torch.empty((lib.vnb_nms3d_workspace_bytes(B, p.npoint),), dtype=torch.uint8, device=dev)
        # the reference's three outputs (model.py:135-137) over this rank's batch; first rec.nms_count rows are valid
        s.bboxes_pred, s.class_scores_pred = E((B * p.npoint, 8, 3)), E((B * p.npoint, NC))
        s.batch_idx = E((B * p.npoint,), i32)
        s.rec.bboxes_pred, s.rec.class_scores_pred, s.rec.batch_idx = s.bboxes_pred, s.class_scores_pred, s.batch_idx
        mmax = max(max(sa.npoint for sa in cfg.sa), cfg.proposal.npoint)
        s.sa_ws = torch.empty((lib.vnb_sa_workspace_bytes(B, mmax, 64),), dtype=torch.uint8, device=dev)
        s.fps_tie = torch.zeros((B,), dtype=i32, device=dev)   # rounds of the identity prefix proven at the sa2 level, per cloud
        s.fps_ws = torch.empty((lib.vnb_fps_nested_workspace_bytes(B, max(sa.npoint for sa in cfg.sa)),), dtype=torch.uint8, device=dev)
        s.done = torch.cuda.Event()
        s.samp = torch.cuda.Stream(device=dev)   # sampling chain (FPS + gathers)
        s.aux = torch.cuda.Stream(device=dev)    # neighbour searches (ball query, three_nn)
        s.graph = None
        s.used = False
        return s

    # ------------------------------------------------------------------------------------------------ launches
    def _linear(self, rows, x, layer, act, out32, out16, st, residual=None):
        check(lib.vnb_linear(rows, layer.cin, layer.cout, dptr(x), dptr(layer.W),
                             dptr(layer.img) if self.precision == PRECISION_TENSOR else None, dptr(layer.b),
                             dptr(residual), 1 if act else 0, dptr(out32), dptr(out16), int(self.precision), _sp(st)))

    def _sa(self, li, xyz, feat, n, c, new_xyz, idx, m, q, out, st, ws=None, cnt=None):
        l1, l2, l3, lf, _ = self.sa_layers[li]
        B = self.B
        tc = self.precision == PRECISION_TENSOR
        if lf is not None:  # hoisted layer 1: q = feat @ W1[3:] + b1, fp16, once per source point
            self._linear(B * n, feat, lf, False, None, q, st)
        check(lib.vnb_sa_group_mlp_max_counted(B, n, c, m, 64, dptr(xyz), dptr(feat), dptr(new_xyz), dptr(idx), dptr(cnt),
                                               l1.cout, l2.cout, l3.cout, dptr(l1.W), dptr(l1.b), dptr(l2.W), dptr(l2.b),
                                               dptr(l3.W), dptr(l3.b), dptr(l1.img) if (tc and lf is None) else None,
                                               dptr(l2.img) if tc else None, dptr(l3.img) if tc else None,
                                               dptr(q) if lf is not None else None, dptr(out), int(self.precision),
                                               dptr(ws), _sp(st)))

    def _enqueue(self, s, main):
        """Enqueue one forward over slot `s`: sampling chain on s_samp, neighbour searches on s_aux, features on main."""
        cfg, B = self.cfg, self.B
        samp, aux = s.samp, s.aux
        ev = torch.cuda.Event
        tl = self.timeline is not None

        def mark(name, stream):
            if tl:
                e = torch.cuda.Event(enable_timing=True)
                e.record(stream)
                self.timeline.append((self._step, name, e))
        mark("start", main)
        e_in = ev(); e_in.record(main)
        samp.wait_event(e_in); aux.wait_event(e_in)
        # ---- sampling chain (xyz only): FPS -> gather, level after level; then the proposal FPS on the seeds
        e_lv = []
        src = s.xyz
        for li, l in enumerate(s.lv):
            if li == 0:   # the only real search (raw cloud); deeper levels sample an FPS-ordered set
                if not (self.debug_skip_fps1 and s.used):
                    check(lib.vnb_farthest_point_sample(B, l.n, l.m, dptr(src), dptr(l.fps), _sp(samp)))
            elif li == 1:  # src = sa1's picks in order: PROVE the identity prefix once, keep what was proven per cloud
                check(lib.vnb_farthest_point_sample_nested_proof(B, l.n, l.m, dptr(src), dptr(l.fps), dptr(s.fps_ws),
                                                                 dptr(s.fps_tie), _sp(samp)))
            else:  # src = a prefix of the proven set: covered by that proof (include/votenet_b200.h), no check needed
                check(lib.vnb_farthest_point_sample_nested_hint(B, l.n, l.m, dptr(src), dptr(l.fps), dptr(s.fps_ws),
                                                                dptr(s.fps_tie), _sp(samp)))
            check(lib.vnb_gather_point(B, l.n, l.m, dptr(src), dptr(l.fps), dptr(l.xyz), _sp(samp)))
            e = ev(); e.record(samp); e_lv.append(e)
            mark(f"fps{li + 1}", samp)
            src = l.xyz
        seeds_xyz = s.lv[1].xyz
        p = cfg.proposal
        check(lib.vnb_farthest_point_sample_nested_hint(B, s.lv[1].m, p.npoint, dptr(seeds_xyz), dptr(s.p_fps), dptr(s.fps_ws),
                                                        dptr(s.fps_tie), _sp(samp)))
        e_pf = ev(); e_pf.record(samp)
        mark("fps_prop", samp)
        # ---- neighbour searches (xyz + centroids only)
        e_bq = []
        src = s.xyz
        # the raw cloud's search grid needs xyz only: it is built while the sa1 FPS runs
        check(lib.vnb_query_ball_point_prepare(B, s.lv[0].n, float(cfg.sa[0].radius), dptr(src), dptr(s.lv[0].bq_ws), _sp(aux)))
        for li, l in enumerate(s.lv):
            aux.wait_event(e_lv[li])
            if li > 0:
                check(lib.vnb_query_ball_point_prepare(B, l.n, float(cfg.sa[li].radius), dptr(src), dptr(l.bq_ws), _sp(aux)))
            check(lib.vnb_query_ball_point_prepared(B, l.n, l.m, float(cfg.sa[li].radius), 64, dptr(src), dptr(l.xyz),
                                                    dptr(l.idx), dptr(l.cnt), dptr(l.bq_ws), _sp(aux)))
            e = ev(); e.record(aux); e_bq.append(e)
            mark(f"bq{li + 1}", aux)
            src = l.xyz
        for f, (u, kx) in zip(s.fp, ((s.lv[2].xyz, s.lv[3].xyz), (s.lv[1].xyz, s.lv[2].xyz))):
            check(lib.vnb_three_nn(B, f.n, f.m, dptr(u), dptr(kx), dptr(f.dist), dptr(f.idx), _sp(aux)))
        e_nn = ev(); e_nn.record(aux)
        # ---- feature chain
        src_xyz, src_feat, c = s.xyz, s.feat, cfg.feature_dim
        for li, l in enumerate(s.lv):
            main.wait_event(e_bq[li])
            mark(f"sa{li + 1}_begin", main)
            self._sa(li, src_xyz, src_feat, l.n, c, l.xyz, l.idx, l.m, l.q, l.feat, main, s.sa_ws, l.cnt)
            mark(f"sa{li + 1}", main)
            src_xyz, src_feat, c = l.xyz, l.feat, cfg.sa[li].mlp[-1]
        main.wait_event(e_nn)
        pts2 = s.lv[3].feat
        ns = s.lv[1].m
        cf = cfg.seed_feat_dim
        if self.fuse_fp:
            # fp1: interpolate + concat + 2 layers in one kernel; fp2 + the voting stack in one kernel (csrc/fp_chain.cu)
            f1, f2 = s.fp
            fp_module_fused(f1.dist, f1.idx, s.lv[2].feat, pts2, [self.store.layer(f"fp1/conv_{i}") for i in range(2)],
                            f1.h[-1], stream=main)
            mark("fp", main)
            # (the seed features are consumed inside the kernel only — the vote residual stays on chip — so they are not stored)
            fp_module_fused(f2.dist, f2.idx, s.lv[1].feat, f1.h[-1].view(B, f1.n, -1),
                            [self.store.layer(f"fp2/conv_{i}") for i in range(2)], None,
                            vote=(self.vote_fused, self.vote_x0, seeds_xyz, s.votes_xyz, s.votes_feat), stream=main)
        else:
            for fi, (f, skip, scope) in enumerate(zip(s.fp, (s.lv[2].feat, s.lv[1].feat), ("fp1", "fp2"))):
                check(lib.vnb_fp_interpolate_concat(B, f.n, f.m, f.c1, f.c2, dptr(f.dist), dptr(f.idx), dptr(skip), dptr(pts2),
                                                    dptr(f.cat), _sp(main)))
                x = f.cat
                for i in range(len(cfg.fp_mlp)):
                    self._linear(B * f.n, x, self.store.layer(f"{scope}/conv_{i}"), True, f.h[i], None, main)
                    x = f.h[i]
                pts2 = x
            mark("fp", main)
            check(lib.vnb_concat2(B * ns, 3, cf, dptr(seeds_xyz), dptr(pts2), dptr(s.seeds), _sp(main)))
            x = s.seeds
            nv = len(cfg.vote_units)
            for i in range(nv):
                self._linear(B * ns, x, self.store.layer(f"voting{i}"), i < nv - 1, s.vh[i], None, main,
                             residual=s.seeds if i == nv - 1 else None)
                x = s.vh[i]
            check(lib.vnb_split2(B * ns, 3, cf, dptr(x), dptr(s.votes_xyz), dptr(s.votes_feat), _sp(main)))
        mark("vote", main)
        main.wait_event(e_pf)
        check(lib.vnb_gather_point(B, ns, p.npoint, dptr(s.votes_xyz), dptr(s.p_fps), dptr(s.p_xyz), _sp(main)))
        check(lib.vnb_query_ball_point(B, ns, p.npoint, float(p.radius), 64, dptr(s.votes_xyz), dptr(s.p_xyz),
                                       dptr(s.p_idx), dptr(s.p_cnt), _sp(main)))
        self._sa(len(cfg.sa), s.votes_xyz, s.votes_feat, ns, cf, s.p_xyz, s.p_idx, p.npoint, s.p_q, s.p_feat, main, s.sa_ws,
                 s.p_cnt)
        x = s.p_feat
        for i in range(len(p.mlp2)):
            self._linear(B * p.npoint, x, self.store.layer(f"proposal/conv_post_{i}"), i < len(p.mlp2) - 1, s.p_h[i], None,
                         main)
            x = s.p_h[i]
        mark("proposal", main)
        r = s.rec
        check(lib.vnb_decode_nms3d(B, p.npoint, dptr(s.p_xyz), dptr(x), dptr(self.mean_size), float(cfg.nms_iou),
                                   dptr(r.bboxes), dptr(r.scores), dptr(r.objectness), dptr(r.class_scores), dptr(r.keep),
                                   dptr(r.nms_idx), dptr(r.nms_key), dptr(r.nms_count), dptr(s.bboxes_pred),
                                   dptr(s.class_scores_pred), dptr(s.batch_idx), dptr(s.nms_ws), _sp(main)))
        mark("nms", main)
        # join the side streams back (all their work has been consumed through events; this keeps capture well-formed)
        e1, e2 = ev(), ev()
        e1.record(samp); e2.record(aux)
        main.wait_event(e1); main.wait_event(e2)

    # ------------------------------------------------------------------------------------------------ public API
    def infer_device(self, xyz, feat, stream=None):
        """xyz (B,N,3), feat (B,N,C) CUDA f32 -> DetectionRecord (device views; valid until this slot is reused, i.e.
        until `slots` calls later).  Besides the wire fields it carries the reference's three outputs over this rank's
        batch (model.py:135-137): rec.bboxes_pred (B*K,8,3), rec.class_scores_pred (B*K,10), rec.batch_idx (B*K,), first
        rec.nms_count rows valid.  Asynchronous on `stream` (default: current stream)."""
        with torch.cuda.device(self.device):
            main = stream if stream is not None else torch.cuda.current_stream(self.device)
            s = self._next_slot(main)
            with torch.cuda.stream(main):
                s.xyz.copy_(xyz, non_blocking=True)
                s.feat.copy_(feat, non_blocking=True)
            self._run_slot(s, main)
        return s.rec

    def _next_slot(self, main):
        """Round-robin slot; its previous forward may have run on another stream, so `main` waits for it first."""
        s = self.slots[self._step % len(self.slots)]
        self._step += 1
        if s.used:
            main.wait_event(s.done)
        return s

    def _run_slot(self, s, main):
        if not self.use_graph:
            n0 = lib.vnb_launch_count()
            self._enqueue(s, main)
            self.launches_per_forward = int(lib.vnb_launch_count() - n0)
            s.used = True
        else:
            if s.graph is None:
                # warm-up outside capture (sets function attributes, packs nothing new), then capture
                self._enqueue(s, main)
                s.used = True
                torch.cuda.synchronize(self.device)
                g = torch.cuda.CUDAGraph()
                cap = torch.cuda.Stream(device=self.device)
                cap.wait_stream(main)
                n0 = lib.vnb_launch_count()
                with torch.cuda.graph(g, stream=cap, capture_error_mode="thread_local"):
                    self._enqueue(s, torch.cuda.current_stream(self.device))
                self.launches_per_forward = int(lib.vnb_launch_count() - n0)
                main.wait_stream(cap)
                s.graph = g
            with torch.cuda.stream(main):
                s.graph.replay()
        s.done.record(main)

    def infer_host(self, xyz_pinned, feat_pinned, out_pinned, stream=None):
        """End-to-end call with HOST buffers: pinned inputs are copied host->device, the forward runs, and the detection
        record is copied device->host into `out_pinned` (uint8, rec.nbytes).  Asynchronous; returns the slot's event."""
        with torch.cuda.device(self.device):
            main = stream if stream is not None else torch.cuda.current_stream(self.device)
            s = self._next_slot(main)
            with torch.cuda.stream(main):
                s.xyz.copy_(xyz_pinned, non_blocking=True)
                s.feat.copy_(feat_pinned, non_blocking=True)
            self._run_slot(s, main)
            with torch.cuda.stream(main):
                out_pinned.copy_(s.rec.buf, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(main)
            s.done.record(main)   # the slot is busy until its record has left for the host
        return ev

    @property
    def record_nbytes(self):
        return self.slots[0].rec.nbytes
