"""The other BASELINE.json configurations, measured in the same run as bench.py's headline and
reported in the `workloads` block of its JSON line:

  configs[2]  HeightToNormal on a synthetic 8192^2 Gray height map (FAST and EXACT arithmetic);
              N > 1: horizontal strips, the halo row read by the kernel from the neighbour GPU's
              mailbox over NVLink (strong scaling, src/node/height_to_normal.rs:16-77)
  configs[3]  Resize Lanczos3 and Gaussian 1024^2 -> 8192^2 RGBA (src/shared.rs:141-216, the idiom of
              tests/integration_tests.rs:353-377); N > 1: row strips of the result, source replicated
  configs[4]  a batch of 64 32-node graphs at 4096^2, whole graphs split over the ranks (strong scaling)
  north star  one 8192^2 RGBA Mix graph, Pow(Multiply(A,B),B) (weak scaling: one per rank)

Every entry carries ms, algorithmic bytes, the roofline fraction and a parity sample against the CPU
oracle.  All device times are CUDA events on the library's stream, max over ranks.

Collective discipline (N > 1): a workload performs the same collectives on every rank whatever
happens -- anything that can fail runs inside `guarded` sections and the failure travels with the
max-reduce that follows, so one rank's exception cannot leave the others waiting.
"""
import ctypes as C
import time

import numpy as np


class Env:
    def __init__(self, kc, tp, rank, world, barrier, max_over_ranks, peak, all_gather_object=None):
        from kanter_core_b200._lib import call
        self.kc, self.tp, self.ctx = kc, tp, tp._ctx._h
        self.rank, self.world = rank, world
        self.barrier, self.max_over_ranks, self.all_gather_object = barrier, max_over_ranks, all_gather_object
        self.peak = peak
        self.call = call
        self.ev = [C.c_void_p(), C.c_void_p()]
        for e in self.ev:
            call("kc_event_create", C.byref(e))

    def timed(self, fn, steps):
        """ms for `steps` calls of fn, by CUDA events on the library's stream, bracketed by barrier + synchronize;
        returns (max over ranks, this rank's)."""
        # two untimed calls with the timed loop's own ownership pattern (the previous result stays alive while the next is
        # computed): whatever the plane pool has to grow for that pattern, it grows here and not inside the timed region
        keep = fn()
        keep = fn()
        self.barrier()
        self.tp.synchronize()
        self.call("kc_event_record", self.ctx, self.ev[0])
        for _ in range(steps):
            keep = fn()
        self.call("kc_event_record", self.ctx, self.ev[1])
        self.tp.synchronize()
        self.barrier()
        ms = C.c_float()
        self.call("kc_event_elapsed_ms", self.ev[0], self.ev[1], C.byref(ms))
        del keep
        return self.max_over_ranks(float(ms.value)), float(ms.value)

    def kernel_times(self, fn, reps=3):
        """device time per kernel kind of `reps` more calls (events around every launch; this rank)"""
        names = {0: "fused_elementwise", 1: "fill", 2: "from_u8", 3: "height_to_normal", 4: "resize_long_window_v", 5: "resize"}
        ms, n = C.c_double(), C.c_uint64()
        self.tp.synchronize()
        self.call("kc_context_set_timing", self.ctx, 1)
        self.call("kc_context_timing_read", self.ctx, -1, C.byref(ms), C.byref(n))
        keep = None
        for _ in range(reps):
            keep = fn()
        self.tp.synchronize()
        out = {}
        for kind, name in names.items():
            self.call("kc_context_timing_read", self.ctx, kind, C.byref(ms), C.byref(n))
            if n.value:
                out[name] = {"ms": ms.value / reps, "launches": n.value / reps}
        self.call("kc_context_set_timing", self.ctx, 0)
        del keep
        return out

    def rows(self, img, plane, y0, n):
        """rows [y0, y0+n) of one plane of a device image, on the host"""
        w = img.size().width
        dst = self.kc.empty_gray(self.tp, w, n)
        self.call("kc_plane_copy_rows", self.ctx, dst._im.planes[0], 0, img._im.planes[plane], int(y0), int(n))
        return dst.planes()[0]

    def roof(self, alg_bytes, ms):
        gbs = alg_bytes / (ms / 1e3) / 1e9 if ms > 0 else 0.0
        return {"bound": "hbm", "achieved": gbs, "peak": self.peak, "unit": "GB/s", "frac": gbs / self.peak, "frac_of_nominal_8TBs": gbs / 8000.0}


def within(got, want):
    w = want.astype(np.float64)
    return int((np.abs(got.astype(np.float64) - w) > 1e-6 + 1e-5 * np.abs(w)).sum())


def bits_equal(a, b):
    return bool(np.array_equal(np.ascontiguousarray(a).view(np.uint32), np.ascontiguousarray(b).view(np.uint32)))


def guarded(env, fn):
    """Run fn(); (result, error string or None).  Never raises."""
    try:
        return fn(), None
    except Exception as ex:  # noqa: BLE001
        return None, repr(ex)[:300]


def agree(env, err):
    """True when no rank failed (one max-reduce on every rank)."""
    return env.max_over_ranks(1.0 if err else 0.0) == 0.0


def strip_rows(height, rank, world):
    base, extra = divmod(height, world)
    y0 = rank * base + min(rank, extra)
    return y0, y0 + base + (1 if rank < extra else 0)


# ---------------------------------------------------------------------------------------------
# configs[2]: HeightToNormal 8192^2
# ---------------------------------------------------------------------------------------------
def wl_height_to_normal(env, steps, size=8192):
    kc, tp = env.kc, env.tp
    H = W = size
    y0, y1 = strip_rows(H, env.rank, env.world)
    st = {}

    def setup():
        import oracle
        full = np.random.default_rng(3).random((H, W), dtype=np.float32)     # the same map on every rank
        st["strip"] = kc.SlotImage.from_planes(tp, [full[y0:y1]])
        if env.rank == 0:   # parity samples: rows 0..7 (they need the WRAPPED halo: the image's last row) and 8 rows mid-strip
            st["want_top"] = oracle.height_to_normal_strip(full[0:8], H, full[H - 1])
            m = (y1 - y0) // 2
            st["mid"] = m
            st["want_mid"] = oracle.height_to_normal_strip(full[m:m + 8], H, full[m - 1])
        if env.world > 1:
            st["outbox"] = kc.HaloLink.outbox(tp, W)
        return True

    _, err = guarded(env, setup)
    # the mailbox handles travel once (object all-gather on every rank, failed or not)
    if env.world > 1:
        handles = env.all_gather_object(st["outbox"].handle() if "outbox" in st else None)
        if err is None and all(h is not None for h in handles):
            _, err = guarded(env, lambda: st.__setitem__("inbox", kc.HaloLink.open(tp, handles[(env.rank - 1) % env.world], W)))
        elif err is None:
            err = "a neighbour has no mailbox"
    if not agree(env, err):
        return {"unavailable": err or "another rank failed"}
    counter = [0]

    def step():
        if env.world == 1:
            return kc.height_to_normal(tp, st["strip"])
        counter[0] += 1     # ONE launch: the stencil kernel publishes my last row, reads the neighbour's and acknowledges it
        return kc.height_to_normal_strip_exchange(tp, st["strip"], st["outbox"], st["inbox"], counter[0], H)

    out = {"workload": "configs[2]: HeightToNormal %dx%d Gray -> RGBA%s" % (W, H, "" if env.world == 1 else ", %d horizontal strips; one launch per strip and step: the stencil kernel publishes its last row, reads the row above from the neighbour's mailbox over NVLink (CUDA IPC) and acknowledges it; no collective" % env.world),
           "scaling": "strong" if env.world > 1 else "single GPU", "pixels": H * W,
           "algorithmic_bytes": H * W * 16, "algorithmic_bytes_note": "4 B read + 12 B written per pixel; the alpha plane stays a constant descriptor (20 B/px if a caller insists on alpha pixels)",
           "reference": "src/node/height_to_normal.rs:16-77"}
    for mode, mode_id in (("fast", kc.MATH_FAST), ("exact", kc.MATH_EXACT)):
        def run():
            tp.set_math_mode(mode_id)
            res = None
            for _ in range(3):
                res = step()
            return res
        res, err = guarded(env, run)
        del res        # the timed loop keeps one result alive while the next is computed, exactly like the warm-up: no third set of planes
        ms_all, _ = env.timed(step, steps) if agree(env, err) else (0.0, 0.0)
        if ms_all <= 0.0:
            out[mode] = {"unavailable": err or "another rank failed"}
            continue
        ms = ms_all / steps
        entry = {"ms": ms, "mpixel_per_s": H * W / 1e6 / (ms / 1e3), "roofline": env.roof((y1 - y0) * W * 16, ms)}
        last, lerr = guarded(env, step)            # EVERY rank takes this step (the mailbox protocol counts steps); rank 0 checks its rows
        if env.rank == 0:
            def parity():
                res = last
                if res is None:
                    raise RuntimeError(lerr)
                bad = 0
                exact = True
                for c in range(3):
                    for want, r0 in ((st["want_top"][c], 0), (st["want_mid"][c], st["mid"])):
                        got = env.rows(res, c, r0, 8)
                        bad += within(got, want)
                        exact &= bits_equal(got, want)
                return bad, exact
            (pr, perr) = guarded(env, parity)
            if perr is None:
                bad, exact = pr
                ok = exact if mode == "exact" else bad == 0
                entry["parity"] = {"ok": bool(ok), "bit_exact": exact, "samples_outside_1e-5rel_1e-6abs": bad, "sample": "rows 0..7 (wrapped halo) and 8 rows mid-strip x 3 planes vs CPU oracle"}
            else:
                entry["parity"] = {"ok": False, "error": perr}
        del last
        out[mode] = entry
    tp.set_math_mode(kc.MATH_FAST)
    if env.world > 1:
        t, _ = guarded(env, lambda: kc.halo_timeouts(tp))
        out["halo_wait_timeouts"] = t
        out["halo_bytes_per_boundary"] = W * 4
        env.barrier()                  # nobody unmaps a mailbox a neighbour may still be reading
        for k in ("inbox", "outbox"):
            if k in st:
                st[k].close()
    return out


# ---------------------------------------------------------------------------------------------
# configs[3]: Resize 1024^2 -> 8192^2 RGBA
# ---------------------------------------------------------------------------------------------
def wl_resize(env, steps, src=1024, dst=8192):
    kc, tp = env.kc, env.tp
    from kanter_core_b200 import ResizeFilter
    from kanter_core_b200._lib import kc_image
    S, D = src, dst
    y0, y1 = strip_rows(D, env.rank, env.world)
    st = {}

    def setup():
        r = np.random.default_rng(4)
        st["planes"] = [r.random((S, S), dtype=np.float32) for _ in range(4)]
        st["img"] = kc.SlotImage.from_planes(tp, st["planes"])
        return True

    _, err = guarded(env, setup)
    if not agree(env, err):
        return {"unavailable": err or "another rank failed"}
    alg = 4 * (S * S * 4 + D * D * 4)
    out = {"workload": "configs[3]: Resize %dx%d -> %dx%d RGBA (4 planes)%s" % (S, S, D, D, "" if env.world == 1 else ", %d row strips of the result, source replicated, no inter-GPU traffic" % env.world),
           "scaling": "strong" if env.world > 1 else "single GPU", "pixels": D * D, "algorithmic_bytes": alg,
           "reference": "src/shared.rs:141-216 -> image 0.24.0 imageops::resize; idiom tests/integration_tests.rs:353-377"}
    for filt in (ResizeFilter.Lanczos3, ResizeFilter.Gaussian):
        def step(filt=filt):
            o = kc_image()
            if env.world == 1:
                env.call("kc_resize", env.ctx, C.byref(st["img"]._im), D, D, int(filt), C.byref(o))
            else:
                env.call("kc_resize_rows", env.ctx, C.byref(st["img"]._im), D, D, int(filt), y0, y1 - y0, C.byref(o))
            return kc.SlotImage(tp._ctx, o)

        def warm():
            for _ in range(3):
                step()
        _, err = guarded(env, warm)
        ms_all, _ = env.timed(step, steps) if agree(env, err) else (0.0, 0.0)
        if ms_all <= 0.0:
            out[filt.name.lower()] = {"unavailable": err or "another rank failed"}
            continue
        ms = ms_all / steps
        entry = {"ms": ms, "mpixel_per_s": D * D / 1e6 / (ms / 1e3), "math": "fast",
                 "roofline": env.roof(4 * (S * S * 4 + (y1 - y0) * D * 4), ms)}
        if env.rank == 0:
            def parity():
                import oracle
                res = step()
                want = oracle.resize_plane(st["planes"][1], D, D, int(filt))         # one whole plane on the CPU: a few seconds
                bad = 0
                for r0 in (0, (y1 - y0) // 2, (y1 - y0) - 16):
                    got = env.rows(res, 1, r0, 16)
                    bad += within(got, want[y0 + r0:y0 + r0 + 16])
                return bad
            bad, perr = guarded(env, parity)
            entry["parity"] = {"ok": perr is None and bad == 0, "samples_outside_1e-5rel_1e-6abs": bad, "error": perr,
                               "sample": "48 rows (top, middle, bottom of the strip) of plane G vs the CPU oracle's image-0.24 restatement"}
        out[filt.name.lower()] = entry
    return out


# ---------------------------------------------------------------------------------------------
# configs[4]: 64 x 32-node graphs at 4096^2, whole graphs split over the ranks
# ---------------------------------------------------------------------------------------------
def wl_graph_batch(env, n_graphs=64, size=4096, distinct=3, replay=True, lanes=3):
    kc, tp = env.kc, env.tp
    from kanter_core_b200 import SlotId
    from tests import graphs
    S = size
    mine = list(range(env.rank, n_graphs, env.world))
    st = {"sets": []}

    def setup():
        g, out = graphs.config5_graph(S)
        st["g"], st["out"] = g, out
        for d in range(max(1, distinct)):
            inputs = graphs.config5_inputs(100 + env.rank * 1000 + d, S)
            lg = tp.new_live_graph()
            lg.set_node_graph(g)
            lg.set_replay(replay)       # evaluation replay: the batch re-evaluates ONE graph structure on new inputs, the case it is for
            imgs = [kc.SlotImage.from_planes(tp, planes) for planes in inputs]
            for eid, img in enumerate(imgs):
                lg.embed_slot_data_with_id(kc.SlotData.new(0, 0, img), eid)
            st["sets"].append((lg, imgs, inputs if (env.rank == 0 and d == 0) else None))
        return True

    def run_share():
        # independent graphs side by side: the reference's thread pool runs ready nodes of the batch concurrently
        # (src/process_pack.rs:27); here the replays of the `distinct` live graphs alternate between `lanes` side streams
        with tp.concurrent(lanes if replay else 1):
            for i, _gid in enumerate(mine):
                lg, imgs, _ = st["sets"][i % len(st["sets"])]
                for eid, img in enumerate(imgs):       # "new inputs arrived": everything downstream is dirty again
                    lg.replace_embedded(img, eid)
                lg.request(st["out"])

    def warm():
        setup()
        # Hot tapes are specialised in the background once they have been seen three times, a bounded number at a time, and
        # a kernel that arrives makes every plan re-capture: passes and waits alternate until nothing is compiling any more,
        # so that the timed pass measures the kernels (and the replays) that serve the batch from then on -- not a compile
        # that happened to land in the middle of it (seen once: 0.56 instead of 0.33 ms per graph)
        for _ in range(4):
            run_share()
            kc.jit_wait()
        run_share()
        run_share()                                # (with replay: ordinary pass, capture, and from here on replays)
        tp.synchronize()
        return True

    _, err = guarded(env, warm)
    ms_all, _ = env.timed(run_share, 1) if agree(env, err) else (0.0, 0.0)
    if ms_all <= 0.0:
        return {"unavailable": err or "another rank failed"}
    lg0 = st["sets"][0][0]
    stats = lg0.last_run_stats()
    per_graph = ms_all / len(mine)
    mpix = S * S / 1e6
    out = {"workload": "configs[4]: %d x 32-node graphs (Separate/Mix/HeightToNormal/Resize/Combine + nested Graph) at %dx%d%s" % (n_graphs, S, S, "" if env.world == 1 else ", whole graphs split over %d ranks, no collective" % env.world),
           "scaling": "strong" if env.world > 1 else "single GPU", "math": "fast", "graphs": n_graphs, "graphs_per_gpu": len(mine),
           "evaluation_replay": dict(lg0_stats := st["sets"][0][0].replay_stats(), on=replay),
           "concurrent_lanes": lanes if replay else 1,
           "ms_total": ms_all, "ms_per_graph_per_gpu": per_graph, "mpixel_per_s": n_graphs * mpix / (ms_all / 1e3),
           "kernels_per_graph": stats["kernels"], "fused_groups_per_graph": stats["fused_groups"],
           "algorithmic_bytes": stats["algorithmic_bytes"], "algorithmic_bytes_note": "per graph, as the library counted it: each distinct plane read once + each result written once per kernel",
           "roofline": env.roof(stats["algorithmic_bytes"], per_graph),
           "reference": "src/node/node_type.rs:213-267 over src/node/{mix,height_to_normal,separate_rgba,combine_rgba,graph}.rs"}
    if env.rank == 0:
        kt, _ = guarded(env, lambda: env.kernel_times(run_share, 1))
        if kt:
            out["kernel_ms_per_graph"] = {k: {"ms": v["ms"] / len(mine), "launches": v["launches"] / len(mine)} for k, v in kt.items()}

        def parity():
            lg, _imgs, inputs = st["sets"][0]
            t0 = time.perf_counter()
            want = graphs.config5_oracle(st["g"], st["out"], inputs)
            cpu_s = time.perf_counter() - t0
            got = lg.slot_data(st["out"], SlotId(0)).image.planes()
            bad = sum(within(got[c], want[c]) for c in range(4))
            worst = max(float(np.abs(got[c].astype(np.float64) - want[c]).max()) for c in range(4))
            return {"ok": bad == 0, "samples_outside_1e-5rel_1e-6abs": bad, "samples": 4 * S * S, "max_abs_err": worst,
                    "sample": "every sample of graph 0's output vs the CPU oracle", "cpu_oracle_seconds_one_graph_one_thread": cpu_s}
        p, perr = guarded(env, parity)
        out["parity"] = p if perr is None else {"ok": False, "error": perr}
    st["sets"].clear()
    return out


# ---------------------------------------------------------------------------------------------
# north star: one 8192^2 RGBA Mix graph per GPU
# ---------------------------------------------------------------------------------------------
def wl_mix_8192(env, steps, size=8192):
    kc, tp = env.kc, env.tp
    from kanter_core_b200 import MixType, Node, NodeType, SlotId
    S = size
    st = {}

    def setup():
        r = np.random.default_rng(50 + env.rank)
        q = S // 2
        # 8192^2 planes tiled from 4096^2 random quadrants (the values matter, their period does not)
        A = [np.tile(r.random((q, q), dtype=np.float32), (2, 2)) for _ in range(3)]
        B = [np.tile(r.random((q, q), dtype=np.float32), (2, 2)) for _ in range(3)]
        st["A"], st["B"] = [a[:64].copy() for a in A], [b[:64].copy() for b in B]
        one = np.ones((1, 1), np.float32)
        ia = kc.SlotImage.from_planes(tp, A + [np.broadcast_to(one, (S, S))])
        ib = kc.SlotImage.from_planes(tp, B + [np.broadcast_to(one, (S, S))])
        del A, B
        lg = tp.new_live_graph()
        lg.embed_slot_data_with_id(kc.SlotData.new(0, 0, ia), 0)
        lg.embed_slot_data_with_id(kc.SlotData.new(0, 0, ib), 1)
        a = lg.add_node(Node.new(NodeType.Embed(0)))
        b = lg.add_node(Node.new(NodeType.Embed(1)))
        mul = lg.add_node(Node.new(NodeType.Mix(MixType.Multiply)))
        pw = lg.add_node(Node.new(NodeType.Mix(MixType.Pow)))
        o = lg.add_node(Node.new(NodeType.OutputRgba("out")))
        for (x, y, s) in ((a, mul, 0), (b, mul, 1), (mul, pw, 0), (b, pw, 1), (pw, o, 0)):
            lg.connect(x, y, SlotId(0), SlotId(s))
        st.update(lg=lg, ia=ia, ib=ib, out=o)
        return True

    def step():
        st["lg"].replace_embedded(st["ia"], 0)
        st["lg"].replace_embedded(st["ib"], 1)
        st["lg"].request(st["out"])

    def warm():
        setup()
        for _ in range(4):
            step()
        kc.jit_wait()
        step()
        tp.synchronize()
        return True

    _, err = guarded(env, warm)
    ms_all, _ = env.timed(step, steps) if agree(env, err) else (0.0, 0.0)
    if ms_all <= 0.0:
        return {"unavailable": err or "another rank failed"}
    ms = ms_all / steps
    alg = st["lg"].last_run_stats()["algorithmic_bytes"]
    out = {"workload": "north star: Pow(Multiply(A,B),B) on two synthetic %dx%d RGBA f32 images -> OutputRgba, one graph per GPU" % (S, S),
           "scaling": "weak" if env.world > 1 else "single GPU", "math": "fast", "pixels": S * S, "ms": ms,
           "mpixel_per_s": env.world * S * S / 1e6 / (ms / 1e3), "algorithmic_bytes": alg, "roofline": env.roof(alg, ms),
           "reference": "src/node/mix.rs:51-302"}
    if env.rank == 0:
        def parity():
            import oracle
            im = st["lg"].slot_data(st["out"], SlotId(0)).image
            bad = 0
            for c in range(3):
                want = oracle.mix_plane(4, oracle.mix_plane(2, st["A"][c], st["B"][c]), st["B"][c])
                bad += within(env.rows(im, c, 0, 64), want)
            alpha = env.rows(im, 3, 0, 64)
            return bad + int((alpha != 1.0).sum())
        bad, perr = guarded(env, parity)
        out["parity"] = {"ok": perr is None and bad == 0, "samples_outside_1e-5rel_1e-6abs": bad, "error": perr, "sample": "64 rows x 4 planes vs CPU oracle"}
    st.clear()
    return out


# ---------------------------------------------------------------------------------------------
# small-graph latency: the 32-node graph at 256^2, where the HOST is the bound (src/engine.rs:128-307 polls every 1 ms)
# ---------------------------------------------------------------------------------------------
def wl_small_graph(env, size=256, evals=300):
    kc, tp = env.kc, env.tp
    from kanter_core_b200 import SlotId
    from tests import graphs
    out_d = {"workload": "the 32-node graph of configs[4] at %dx%d, one evaluation after another (new inputs embedded each time): microseconds per evaluation, host + device, wall clock over %d evaluations" % (size, size, evals),
             "reference": "src/engine.rs:128-307 (the reference needs >= 1 ms per DAG level: a polling loop)"}

    def measure(replay):
        g, out = graphs.config5_graph(size)
        inputs = graphs.config5_inputs(77, size)
        lg = tp.new_live_graph()
        lg.set_node_graph(g)
        lg.set_replay(replay)
        imgs = [kc.SlotImage.from_planes(tp, planes) for planes in inputs]
        for eid, img in enumerate(imgs):
            lg.embed_slot_data_with_id(kc.SlotData.new(0, 0, img), eid)

        def one():
            for eid, img in enumerate(imgs):
                lg.replace_embedded(img, eid)
            lg.request(out)
        for _ in range(20):
            one()
        tp.synchronize()
        t0 = time.perf_counter()
        for _ in range(evals):
            one()
        tp.synchronize()
        us = (time.perf_counter() - t0) / evals * 1e6
        got = lg.slot_data(out, SlotId(0)).image.planes()
        want = graphs.config5_oracle(g, out, inputs)
        ok = all(bits_equal(got[c], want[c]) for c in range(4)) if tp._ctx is not None else False
        return us, lg.last_run_stats()["kernels"], lg.replay_stats(), ok

    def body():
        tp.set_math_mode(kc.MATH_EXACT)
        a = measure(False)
        b = measure(True)
        tp.set_math_mode(kc.MATH_FAST)
        return a, b
    res, err = guarded(env, body)
    if err is not None:
        tp.set_math_mode(kc.MATH_FAST)
        out_d["unavailable"] = err
        return out_d
    (us0, k0, _, ok0), (us1, k1, st1, ok1) = res
    out_d.update({"math": "exact", "ordinary_us_per_evaluation": us0, "replay_us_per_evaluation": us1, "kernels_per_evaluation": k0,
                  "replay": "evaluation replay (kc_live_graph_set_replay): the captured CUDA graph of the same %d kernels, one launch" % k1,
                  "replay_stats": st1, "parity": {"ok": bool(ok0 and ok1), "sample": "every sample of the last evaluation of either mode, bit-identical to the CPU oracle"}})
    return out_d


def run_all(env, steps):
    """Every workload in turn; each is independent of the others' success."""
    out = {}
    for name, fn in (("height_to_normal_8192", lambda: wl_height_to_normal(env, steps)),
                     ("resize_1024_to_8192_rgba", lambda: wl_resize(env, steps)),
                     ("graphs32_batch64_4096", lambda: wl_graph_batch(env)),
                     ("mix_rgba_8192", lambda: wl_mix_8192(env, steps)),
                     ("graph32_latency_256", lambda: wl_small_graph(env))):
        t0 = time.perf_counter()
        out[name] = fn()
        if isinstance(out[name], dict):
            out[name]["wall_s"] = round(time.perf_counter() - t0, 2)
        env.kc.jit_wait(1)
        env.call("kc_context_trim", env.ctx)
    return out
