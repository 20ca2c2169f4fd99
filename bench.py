#!/usr/bin/env python3
"""bench.py — NDT-D2D registrations/sec on B200 (BASELINE.json metric) + roofline + CPU baseline.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--pairs B] [--impl reference]
  (N > 1: launched by torchrun, one rank per GPU; ranks shard the scan pairs, weak scaling)

One "step" = one pass of the hot path over one batch: B scan pairs of BASELINE config C2 (100k-point Velodyne-like
scans, 0.5 m voxels): build the NDT map of both scans (kernel i), register source onto target from an odometry-like
guess (kernel ii inside the device-resident Newton/More-Thuente loop) and compute the pose covariance — i.e.
NDTFeatureFuserHMT::update's local-map + match + covariance (ndt_feature_fuser_hmt.cpp:195-227,356,399-420) /
updateLinkUsingNDTRegistration (ndt_feature_graph.cpp:260-345), batched.
  value : registrations/s with the scans already resident in HBM, results left in HBM (CUDA events)
  e2e   : the same through the C ABI with HOST buffers (pinned scans H2D + results D2H inside the timed region)
"""
import argparse
import ctypes as C
import json
import os
import statistics
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "NDT-D2D registrations/sec (100k-pt scans, 0.5 m voxels)"
UNIT = "registrations/s"
WORKLOAD = "C2: 3-D NDT-D2D scan-pair registration (map build x2 + match + covariance), 100k-pt synthetic Velodyne-like scans, 0.5 m voxels"
CELL = 0.5


def peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


# ------------------------------------------------------------------ clocks
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device):
        self.p = None
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                                       "-i", str(device)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.p = None

    def stop(self):
        if self.p is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.p.terminate()
        try:
            out, _ = self.p.communicate(timeout=5)
        except Exception:
            self.p.kill()
            out = ""
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in out.strip().splitlines():
            f = [x.strip() for x in line.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for nm, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        busy = [s for s in sm if s > 0.5 * max(sm)] or sm
        return {"sm_mhz": statistics.median(busy), "sm_max_mhz": max(mx), "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------ CPU baseline (oracle port, test infrastructure)
def cpu_registrations(tg, sr, T0s, threads, check_stability=0):
    """Times the oracle (restated reference algorithm) on the given pairs with `threads` host threads (one pair per
    thread at a time, like an OpenMP loop over edges).  Returns (seconds, results).  For the first `check_stability`
    pairs the match is repeated (untimed) with 3 OpenMP partial sums inside derivativesNDT — upstream's N_THREADS
    summation order — and with the initial guess nudged by one ulp, to tell whether the reference algorithm reproduces
    ITSELF on that pair (oracle_py.d2d_is_stable, DESIGN.md "Parity")."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import oracle_py as O
    from concurrent.futures import ThreadPoolExecutor

    O.lib()
    keep = {}

    def one(i):
        maps = []
        for c in (tg[i], sr[i]):
            m = O.OracleMap(CELL)
            m.load_point_cloud(c, -1.0)
            m.compute_cells()
            maps.append(m)
        r = O.d2d_match(maps[0], maps[1], T0s[i])
        cov = np.full((6, 6), 0.0)
        if r.pose_changed:
            _, cov = O.d2d_covariance(maps[0], maps[1], r.pose())
        if i < check_stability:
            keep[i] = maps
        return r.pose(), cov, r.iterations

    t0 = time.perf_counter()
    with ThreadPoolExecutor(max_workers=threads) as ex:
        out = list(ex.map(one, range(len(tg))))
    dt = time.perf_counter() - t0

    def again(i):  # True when the oracle reproduces itself (other summation order, 1-ulp nudges of the initial guess)
        return O.d2d_is_stable(keep[i][0], keep[i][1], T0s[i])

    if check_stability:
        with ThreadPoolExecutor(max_workers=threads) as ex:
            alt = list(ex.map(again, range(min(check_stability, len(tg)))))
        return dt, out, alt
    return dt, out


def host_cores():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def run_reference(args):
    """--impl reference: the reference's CPU path.  perception_oru is not buildable here, so this is the
    oracle port timed on all host cores, each step a bounded sample of the same workload."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from ndt_feature_graph_b200 import synth

    cores = host_cores()
    n = max(4, min(args.pairs, 8 * cores))  # bounded sample of the step: 8 pairs per host thread (keeps the tail of the
    # slowest pair small against the step: the CPU arm should not look slower than it is)
    tg, sr, T0s, Ds = synth.velodyne_batch(n, n_base=min(args.base, n), seed=0)
    for _ in range(args.warmup):
        cpu_registrations(tg[:min(n, cores)], sr[:min(n, cores)], T0s[:min(n, cores)], cores)
    t = 0.0
    for _ in range(args.steps):
        dt, _ = cpu_registrations(tg, sr, T0s, cores)
        t += dt
    v = n * args.steps / t
    line = {
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * t / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic", "config": {"workload": WORKLOAD, "pairs_per_step": n},
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": f"{n} scan pairs per step, one pair per host thread, {cores} threads"},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------ the B200 arm
def run_gpu(args):
    import torch
    import torch.distributed as dist

    import ndt_feature_graph_b200 as N
    from ndt_feature_graph_b200 import api, sharding, synth

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback); use --impl reference for the CPU arm")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    B = args.pairs
    # every rank owns B distinct pairs of one workload (weak scaling: per-GPU work fixed): pairs [rank*B, (rank+1)*B)
    tg, sr, T0s, Ds = synth.velodyne_batch(B, n_base=args.base, seed=0, start=rank * B)
    # `lanes` contexts (one host thread + one stream each, the ABI's "one ndtb_ctx per host thread") take the steps in turn:
    # while one step's last registrations (the few that need hundreds of passes) finish on a few SMs, the next step's
    # kernels fill the rest of the GPU.  Every step does the full work on the same B pairs.
    lanes = max(1, args.lanes)
    streams = [torch.cuda.Stream(dev) for _ in range(lanes)]
    engs_d = [N.Engine(local, stream=st.cuda_stream) for st in streams]
    stream, eng = streams[0], engs_d[0]
    prm = eng.default_params()
    # device-resident scans
    d_t = [torch.from_numpy(c).to(dev) for c in tg]
    d_s = [torch.from_numpy(c).to(dev) for c in sr]
    tp = (C.c_void_p * B)(*[t.data_ptr() for t in d_t])
    sp = (C.c_void_p * B)(*[t.data_ptr() for t in d_s])
    tn = (C.c_int64 * B)(*[c.shape[0] for c in tg])
    sn = (C.c_int64 * B)(*[c.shape[0] for c in sr])
    T0c = np.concatenate([np.ascontiguousarray(T.T).ravel() for T in T0s])
    d_ress = [torch.zeros(B * api.RESULT_DTYPE.itemsize, dtype=torch.uint8, device=dev) for _ in range(lanes)]
    d_covs = [torch.zeros(B * 36, dtype=torch.float64, device=dev) for _ in range(lanes)]
    d_res, d_cov = d_ress[0], d_covs[0]
    in_bytes = 16 * (sum(c.shape[0] for c in tg) + sum(c.shape[0] for c in sr))

    comm_stream = torch.cuda.Stream(dev) if world > 1 else None

    step_sync = lanes > 1  # with several lanes every lane waits for its own step before enqueuing its next one

    def step_device(k=0):
        engs_d[k].register_scans_raw(B, tp, tn, sp, sn, T0c.ctypes.data, CELL, -1.0, prm, True, api.DEVICE, api.DEVICE,
                                     d_ress[k].data_ptr(), d_covs[k].data_ptr())
        if step_sync:
            engs_d[k].synchronize()

    def gather_step(records, ev):
        # the only cross-GPU step: gather of the per-edge result records (NCCL over NVLink), always issued by the main
        # thread in step order (collectives of one process group must be issued in the same order on every rank)
        with torch.cuda.stream(comm_stream):
            comm_stream.wait_event(ev)
            sharding.gather_results(records, world * B, rank, world)

    def run_device_steps(n_steps):
        import threading

        do_gather = world > 1 and not os.environ.get("NDTB_BENCH_NO_GATHER")
        snaps, evs = [None] * n_steps, [torch.cuda.Event() for _ in range(n_steps)]
        done = [threading.Event() for _ in range(n_steps)]

        def worker(k):
            torch.cuda.set_device(local)
            for s_ in range(k, n_steps, lanes):
                step_device(k)
                if do_gather:
                    with torch.cuda.stream(streams[k]):
                        snaps[s_] = d_ress[k].clone()  # lane k's next step overwrites its record buffer
                        evs[s_].record(streams[k])
                done[s_].set()

        th = [threading.Thread(target=worker, args=(k,)) for k in range(lanes)]
        for t_ in th:
            t_.start()
        for s_ in range(n_steps):
            done[s_].wait()
            if do_gather:
                gather_step(snaps[s_], evs[s_])
        for t_ in th:
            t_.join()

    def sync_all():
        for st in streams:
            st.synchronize()
        if comm_stream is not None:
            comm_stream.synchronize()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()

    run_device_steps(max(args.warmup, 2 * lanes))  # every lane twice: kernels loaded, its memory pool grown
    sync_all()
    l0 = sum(e.launch_count for e in engs_d)
    clk = ClockSampler(local)
    ev0 = torch.cuda.Event(enable_timing=True)
    ev1s = [torch.cuda.Event(enable_timing=True) for _ in range(lanes)]
    sync_all()
    ev0.record(stream)
    for st in streams[1:]:
        st.wait_event(ev0)
    run_device_steps(args.steps)
    for k in range(lanes):
        ev1s[k].record(streams[k])
    if comm_stream is not None:
        ev1s.append(torch.cuda.Event(enable_timing=True))
        ev1s[-1].record(comm_stream)
    sync_all()
    ms = max(ev0.elapsed_time(e) for e in ev1s)
    clocks = clk.stop()
    launches = sum(e.launch_count for e in engs_d) - l0
    # roofline of the dominant kernel: its launches timed alone (one lane), CUDA events inside the library
    eng.enable_timing(True)
    eng.match_time()
    eng.build_time()
    for _ in range(max(2, min(args.steps, 3))):
        step_device(0)
    match_ms, match_n = eng.match_time()
    build_ms, build_n = eng.build_time()
    eng.enable_timing(False)
    sync_all()
    solo0, solo1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    solo0.record(stream)
    step_device(0)
    solo1.record(stream)
    sync_all()
    solo_ms = solo0.elapsed_time(solo1)
    if world > 1:
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    res = np.frombuffer(d_res.cpu().numpy().tobytes(), dtype=api.RESULT_DTYPE)
    cov = d_cov.cpu().numpy().reshape(B, 6, 6)
    for k in range(1, lanes):
        if True:
            assert np.array_equal(np.frombuffer(d_ress[k].cpu().numpy().tobytes(), dtype=api.RESULT_DTYPE)["T"], res["T"]), "lanes disagree"

    # ---- e2e: host (pinned) scans -> C ABI -> host results, copies inside the timed region
    h_t = [torch.from_numpy(c).pin_memory() for c in tg]
    h_s = [torch.from_numpy(c).pin_memory() for c in sr]
    htp = (C.c_void_p * B)(*[t.data_ptr() for t in h_t])
    hsp = (C.c_void_p * B)(*[t.data_ptr() for t in h_s])
    h_res = np.zeros(B, api.RESULT_DTYPE)
    h_cov = np.zeros((B, 36))

    # The host-buffer leg runs as a caller with a stream of batches would: `lanes` contexts (one host thread each, the
    # ABI's "one ndtb_ctx per host thread"), consecutive steps alternate between them, so the H2D upload of step s+1
    # overlaps the registration kernels of step s.  Every step still uploads its own scans and reads back its own
    # results; the timed region is the wall time until the last step's results are in host memory.
    lanes_d = lanes
    lanes = max(1, args.e2e_lanes)
    engs = [eng] + [N.Engine(local) for _ in range(lanes - 1)]
    h_ress = [np.zeros(B, api.RESULT_DTYPE) for _ in range(lanes)]
    h_covs = [np.zeros((B, 36)) for _ in range(lanes)]
    h_res, h_cov = h_ress[0], h_covs[0]

    def step_host(k=0):
        engs[k].register_scans_raw(B, htp, tn, hsp, sn, T0c.ctypes.data, CELL, -1.0, prm, True, api.HOST, api.HOST,
                                   h_ress[k].ctypes.data, h_covs[k].ctypes.data)

    def run_host_steps(n_steps):
        if lanes == 1:
            for _ in range(n_steps):
                step_host(0)
            return
        import threading

        def worker(k):
            for s_ in range(k, n_steps, lanes):
                step_host(k)

        th = [threading.Thread(target=worker, args=(k,)) for k in range(lanes)]
        for t_ in th:
            t_.start()
        for t_ in th:
            t_.join()

    e2e_steps = max(lanes, min(args.steps, 6))
    if args.no_e2e:  # profiling runs only (ncu): skip the host-buffer leg
        e2e_steps, e2e_s = 0, float("nan")
        h_res["T"] = res["T"]
    else:
        run_host_steps(2 * lanes)  # warm-up of every lane (twice: its memory pool grown)
        sync_all()
        t0 = time.perf_counter()
        run_host_steps(e2e_steps)
        sync_all()
        e2e_s = time.perf_counter() - t0
        for k in range(1, lanes):
            assert np.array_equal(h_ress[k]["T"], h_res["T"]), "e2e lanes disagree"
    if world > 1:
        t = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t.item())
    assert np.array_equal(h_res["T"], res["T"]), "host-buffer and device-buffer paths disagree"

    if rank == 0:
        hbm, hbm_src = peaks()
        passes = (res["n_hess_passes"] + res["n_grad_passes"]).astype(np.float64) + 1.0  # + the covariance pass
        b_pass = 72.0 * (res["n_src_cells"] + res["n_tgt_cells"]) + 16.0 * res["tgt_table_entries"] + 344.0
        alg_bytes_launch = float(((passes - 1.0) * b_pass).sum())  # match kernel only
        t_launch = 1e-3 * match_ms / max(match_n, 1)
        value = world * B * args.steps / (1e-3 * ms)
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic",
            "config": {"workload": WORKLOAD, "pairs_per_step_per_gpu": B, "base_scenes": args.base, "lanes": lanes_d,
                       "points_per_scan": int(np.mean([c.shape[0] for c in tg])),
                       "gaussian_cells_per_map": int(res["n_tgt_cells"].mean()),
                       "n_neighbours": int(prm.n_neighbours), "delta_score": prm.delta_score,
                       "l2": f"inputs {in_bytes / 1e6:.0f} MB per step per GPU (> 126 MB L2), no flush needed",
                       "iterations_mean": float(res["iterations"].mean()),
                       "passes_mean": float(passes.mean()), "converged_frac": float(res["converged"].mean())},
            "clocks": clocks,
            "e2e": {"value": (world * B * e2e_steps / e2e_s) if e2e_steps else None, "unit": UNIT, "h2d_bytes_per_step": in_bytes + 128 * B,
                    "d2h_bytes_per_step": B * (api.RESULT_DTYPE.itemsize + 288), "steps": e2e_steps,
                    "lanes": lanes},
            "gpu_launches": int(launches),
            "roofline": {"kernel": "match_kernel (device-resident Newton loop around the D2D derivative pass)",
                         "bound": "hbm", "achieved": alg_bytes_launch / t_launch / 1e9, "peak": hbm, "unit": "GB/s",
                         "frac": alg_bytes_launch / t_launch / 1e9 / hbm, "peak_source": hbm_src,
                         # dram__bytes_read+write of the two launches of one 592-pair step, ncu --set full capture
                         # profiles/r01b_match_kernel_ncu_full.csv (GB per step; scales with pairs per step)
                         "traffic": 3.196 * B / 592.0, "traffic_unit": "GB per step (both launches of the kernel)",
                         "algorithmic_gb_per_step": alg_bytes_launch / 1e9,
                         "launch_ms": 1e3 * t_launch, "share_of_step": 1e3 * t_launch / solo_ms, "step_ms_one_lane": solo_ms,
                         "note": "working set is L1/L2 resident; the binding limit is fp64 CUDA-core throughput, see DESIGN.md"},
        }
        # kernel (i), the bandwidth-bound one by design: B_build = 16 P + 80 C per map
        alg_build = float(in_bytes) + 80.0 * float((res["n_src_cells"] + res["n_tgt_cells"]).sum())
        t_build = 1e-3 * build_ms / max(build_n, 1)
        line["roofline_build"] = {"kernel": "map build (13 launches of kernel i per step, 2 maps per pair)", "bound": "hbm",
                                  "achieved": alg_build / t_build / 1e9, "peak": hbm, "unit": "GB/s",
                                  "frac": alg_build / t_build / 1e9 / hbm, "build_ms": 1e3 * t_build,
                                  "note": "latency-bound by the order-exact per-cell chains, see DESIGN.md"}
        if world == 1 and not args.no_cpu:
            cores = host_cores()
            # bounded sample: ~0.6 core-seconds per C2 pair -> 10-30 s of wall time on the box's cores
            ns = min(B, args.cpu_pairs if args.cpu_pairs > 0 else max(16, 24 * cores))
            nchk = min(ns, 64)
            dt, out, alt = cpu_registrations(tg[:ns], sr[:ns], T0s[:ns], cores, check_stability=nchk)
            errs = np.array([synth.pose_error(out[i][0], res["T"][i].reshape(4, 4).T) for i in range(ns)])
            # a pair pins parity only if the reference algorithm reproduces itself under its own (OpenMP) change of
            # summation order; basin-hopping registrations amplify 1-ulp differences to O(1) (DESIGN.md "Parity")
            selfc = np.array(alt, dtype=bool)
            line["cpu_baseline"] = {"value": ns / dt, "unit": UNIT, "cores": cores, "kind": "port",
                                    "sample": f"first {ns} scan pairs of the step (map build x2 + match + covariance), "
                                              f"one pair per host thread, {cores} threads, {dt:.1f} s"}
            line["parity"] = {"pairs_checked": ns, "tolerance": 1e-4,
                              "pairs_within_tol": int((errs < 1e-4).sum()),
                              "self_consistency_checked": int(nchk),
                              "oracle_self_consistent": int(selfc.sum()),
                              "self_consistent_within_tol": int((errs[:nchk][selfc] < 1e-4).sum()),
                              "pose_err_max_self_consistent": float(errs[:nchk][selfc].max()) if selfc.any() else None,
                              "pose_err_median": float(np.median(errs))}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=6)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--pairs", type=int, default=592, help="scan pairs per step per GPU")
    ap.add_argument("--base", type=int, default=8, help="ray-cast base scenes per rank")
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-e2e", action="store_true", help="skip the host-buffer (e2e) leg: profiling runs only")
    ap.add_argument("--lanes", type=int, default=3, help="contexts (host thread + stream each) the device-resident leg alternates its steps between")
    ap.add_argument("--e2e-lanes", type=int, default=3, help="contexts (host threads) the e2e leg alternates its steps between")
    ap.add_argument("--cpu-pairs", type=int, default=0, help="pairs of the step timed on the CPU (0 = auto, ~10-30 s)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
