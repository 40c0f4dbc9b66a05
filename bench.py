#!/usr/bin/env python
"""Benchmark of the Oryon inference hot path on B200 (BASELINE.json metric: image-pairs/sec at 1/2/4/8 GPUs; match-kernel
figures against the measured peaks).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 bench.py --gpus N ...

Workload of the contract line = BASELINE config 3 (the full path; config 4 is the same path on the TOYL split, sharded): the
synthetic NOCS-like 2000-pair split (224x224 RGB, 480x640 depth with a planted rigid motion, 34 distinct prompt sets, seeded
random weights of the reference's architecture, PREDICTED masks), B = 32 pairs per step and GPU, the pair list sharded over the
ranks with ``sharding.shard_pairs`` and the result rows all-gathered once at the end.  One "step" = one
``FPM_Pipeline.test_step`` over a batch of 32 pairs: network (CLIP ViT-L/14@336 + swin_b guidance + fusion + decoder) -> masks
-> nearest-neighbour matching -> draws -> lifting -> PointDSC -> pose rows.

  value         pairs/s with the batches already resident in HBM (CUDA events on the launching stream, barrier + synchronize
                on both sides, max over ranks), K steps
  e2e           the same K steps through the public call with HOST batches: pinned RGB / mask / depth frames and prompt
                STRINGS in, pose rows out; every host<->device copy inside the timed region; plus the final all-gather
  roofline      match_tc_kernel at BASELINE config 2 (the configuration the kernel figure is quoted on), measured in this run
                over its own >= 2 s region; burst and sustained peak both printed, the one that applies chosen by region length
  roofline_config5, matcher_config2 (value / e2e of the matcher alone = round 1's headline), network_gemm, h2d_probe
  cpu_baseline  the oracle port of the reference's whole step on the host cores, bounded sample (rank 0, N = 1 only)
`--impl reference` times that CPU port alone: one step = ONE pair of the same workload, all host threads.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

FULL = dict(B=32, split_pairs=2000, distinct_prompts=34, mask="predicted", distinct_batches=4, head_bias_shift=-1.10)
C2 = dict(B=32, D=128, H=120, W=160)
C5 = dict(B=8, D=256, H=240, W=320)
L2_BYTES = 126e6
BPE_SYNTH = os.path.join(ROOT, "tests", "golden", "bpe_synth_vocab.txt.gz")
TRAFFIC_FILE = os.path.join(ROOT, "profiles", "ncu_traffic.json")


DTYPE_NOTE = {
    3: "f16 split operands (hi+lo, 3 tcgen05 products = f32-equivalent) with f32 accumulation in the network; matcher: f16 candidate "
       "pass + f32 re-score; PointDSC f32",
    2: "network: f16 split operands with f32 accumulation -- hi*hi on f16 + both cross terms as one e5m2 x e4m3 product in the CLIP vision "
       "linear layers (2 tensor-pipe units), 3 f16 products elsewhere; all stages within the 1e-3 gate of the f32 reference; matcher: f16 "
       "candidate pass + f32 re-score; PointDSC f32",
    1: "f16 single product (NOT the parity mode)",
}
GEMM_NOTE = {
    3: "every algorithmic product is issued as 3 fp16 products (hi*hi + lo*hi + hi*lo) to hold the 1e-3 feature-map gate",
    2: "CLIP vision linear layers: hi*hi (f16) + one 8-bit product carrying both cross terms = 2 tensor-pipe units per product; every other "
       "GEMM 3 fp16 products; tensor_pipe_tflops counts the units actually issued (f16-equivalent)",
    1: "one fp16 product per algorithmic product",
}


def workload_config() -> dict:
    """`config` of the JSON line -- the SAME dict in both arms (the driver compares them)."""
    return {
        "workload": "config3: full path on the synthetic NOCS-like 2000-pair split -- 224x224 RGB + 480x640 depth pairs -> CLIP ViT-L/14@336 "
                    "+ swin_b guidance + fusion + decoder -> 32x192x192 feature maps + predicted masks -> NN matching (5000-row source "
                    "subsample, th 0.25) -> 500 correspondences -> lift -> PointDSC -> pose; 34 distinct prompt sets; B=32 pairs per step",
        "pairs_per_step": FULL["B"], "split_pairs": FULL["split_pairs"], "distinct_prompts": FULL["distinct_prompts"], "mask": FULL["mask"],
        "weights": "seeded random weights of the reference's architecture (no checkpoint offline); decoder.head.bias shifted by "
                   f"{FULL['head_bias_shift']} so that the predicted masks cover ~13 % of the feature map (unshifted: 79 %)",
        "l2": "a step streams 117 MB of inputs, ~1.7 GB of packed weights and GBs of activations: far beyond the 126 MB L2, no flush needed",
        "parallelism": "pairs sharded over the ranks (sharding.shard_pairs), no data-path collective, one all_gather of 16-float result rows",
        "pipelining": "test.pipelined: the post-network tail of batch k runs on a second stream under the network pass of batch k+1",
    }


def algorithmic_work(B, D, na, nq, elt=4):
    """SURVEY.md 8(d) / BASELINE.md section 4: per pair FLOPs = 2*Na*Nq*D, bytes = (Na+Nq)*D*s + (Na+Nq) + 8*Na."""
    flops = 2.0 * na * nq * D * B
    bytes_ = ((na + nq) * D * elt + (na + nq) + 8 * na) * B
    return flops, bytes_


class ClockSampler:
    """Samples SM clock / throttle reasons while the timed region runs (NVML, falling back to nvidia-smi)."""

    def __init__(self, index):
        self.index, self.samples, self.reasons, self.max_mhz = index, [], set(), None
        self.power_w, self.util, self.power_limit_w = [], [], None
        self._stop = threading.Event()
        self._thr = None
        self._nvml = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self._nvml = pynvml
            self._h = _nvml_handle(pynvml, index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self._h, pynvml.NVML_CLOCK_SM)
            try:
                self.power_limit_w = pynvml.nvmlDeviceGetEnforcedPowerLimit(self._h) / 1e3
            except Exception:
                pass
        except Exception:
            self._nvml = None

    def _loop(self):
        nv = self._nvml
        names = {}
        if nv is not None:
            for n in ("HwSlowdown", "HwThermalSlowdown", "SwThermalSlowdown", "SwPowerCap", "HwPowerBrakeSlowdown"):
                v = getattr(nv, "nvmlClocksThrottleReason" + n, None) or getattr(nv, "nvmlClocksEventReason" + n, None)
                if v:
                    names[v] = n
        while not self._stop.is_set():
            try:
                if nv is not None:
                    self.samples.append(nv.nvmlDeviceGetClockInfo(self._h, nv.NVML_CLOCK_SM))
                    mask = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self._h)
                    for bit, n in names.items():
                        if mask & bit:
                            self.reasons.add(n)
                    try:   # board power and the share of the last sampling window with a kernel resident: evidence for "power-capped" / "never idle"
                        if os.environ.get("ORYON_BENCH_NO_POWER"):
                            raise RuntimeError("disabled")
                        self.power_w.append(nv.nvmlDeviceGetPowerUsage(self._h) / 1e3)
                        self.util.append(nv.nvmlDeviceGetUtilizationRates(self._h).gpu)
                    except Exception:
                        pass
                else:
                    out = subprocess.run(["nvidia-smi", f"--id={self.index}", "--query-gpu=clocks.sm,clocks.max.sm,"
                                          "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
                                          "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap",
                                          "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
                    f = [x.strip() for x in out.strip().split(",")]
                    self.samples.append(int(f[0]))
                    self.max_mhz = int(f[1])
                    for n, v in zip(("HwSlowdown", "HwThermalSlowdown", "SwThermalSlowdown", "SwPowerCap"), f[2:]):
                        if v == "Active":
                            self.reasons.add(n)
            except Exception:
                pass
            self._stop.wait(0.02 if nv is not None else 0.1)

    def __enter__(self):
        self._thr = threading.Thread(target=self._loop, daemon=True)
        self._thr.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        self._thr.join(timeout=2)

    def summary(self):
        out = {"sm_mhz": (statistics.median(self.samples) if self.samples else None), "sm_max_mhz": self.max_mhz,
               "reasons": sorted(self.reasons), "samples": len(self.samples)}
        if self.power_w:
            out.update(power_w=round(statistics.median(self.power_w), 1), power_limit_w=self.power_limit_w,
                       gpu_util_pct=statistics.median(self.util), gpu_util_min_pct=min(self.util))
        return out


def _nvml_handle(pynvml, cuda_index):
    """NVML handle of CUDA device `cuda_index` (by UUID: CUDA_VISIBLE_DEVICES may renumber the devices)."""
    try:
        uuid = str(torch.cuda.get_device_properties(cuda_index).uuid)
        return pynvml.nvmlDeviceGetHandleByUUID(("GPU-" + uuid) if not uuid.startswith("GPU-") else uuid)
    except Exception:
        return pynvml.nvmlDeviceGetHandleByIndex(cuda_index)


def bind_rank_to_gpu_locality(local: int, world: int) -> dict:
    """Pins this rank's threads to the CPUs NVML reports as local to its GPU (same NUMA node / PCIe root), and splits that set
    evenly among the ranks that share it, BEFORE any pinned buffer is allocated (first touch then places the staging memory on
    the GPU's node).  Round 1's records show every GPU of the pool's boxes reporting the same CPU set (one NUMA node): the split
    then at least keeps eight ranks' copy / launch threads off each other's cores.  Returns what was done, for the JSON line."""
    info = {"applied": False}
    try:
        import pynvml
        pynvml.nvmlInit()
        allowed = sorted(os.sched_getaffinity(0))
        words = (max(allowed) // 64) + 1

        def cpus_of(i):
            mask = pynvml.nvmlDeviceGetCpuAffinity(_nvml_handle(pynvml, i), words)
            return tuple(c for c in allowed if (mask[c // 64] >> (c % 64)) & 1)

        sets = [cpus_of(i) for i in range(world)]
        mine = sets[local] or tuple(allowed)
        sharers = [i for i in range(world) if (sets[i] or tuple(allowed)) == mine]
        k, n = sharers.index(local), len(sharers)
        per = max(1, len(mine) // n)
        part = mine[k * per:(k + 1) * per] if k * per < len(mine) else mine
        os.sched_setaffinity(0, set(part))
        info = {"applied": True, "gpu_cpu_set": f"{mine[0]}-{mine[-1]} ({len(mine)})", "ranks_sharing_it": n, "this_rank_cpus": f"{part[0]}-{part[-1]} ({len(part)})"}
        try:
            info["numa_node"] = int(pynvml.nvmlDeviceGetNumaNodeId(_nvml_handle(pynvml, local)))
        except Exception:
            pass
    except Exception as e:
        info["error"] = repr(e)[:200]
    return info


def h2d_probe(dev, world, barrier, mb=256, reps=6) -> dict:
    """Host-to-device bandwidth of this rank's pinned memory with ALL ranks copying at once: what the box gives N concurrent
    uploaders (the ceiling of any e2e figure that is bound by its inputs)."""
    import torch.distributed as dist
    host = torch.empty(mb << 20, dtype=torch.uint8).pin_memory()
    devb = torch.empty(mb << 20, dtype=torch.uint8, device=dev)
    devb.copy_(host, non_blocking=True)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        devb.copy_(host, non_blocking=True)
    e1.record()
    barrier()
    gbs = reps * (mb << 20) / (e0.elapsed_time(e1) * 1e-3) / 1e9
    t = torch.tensor([gbs], device=dev, dtype=torch.float64)
    if world > 1:
        allr = [torch.empty_like(t) for _ in range(world)]
        dist.all_gather(allr, t)
        per = [round(float(x.item()), 1) for x in allr]
    else:
        per = [round(gbs, 1)]
    return {"per_rank_gbs": per, "aggregate_gbs": round(sum(per), 1), "buffer_mb": mb, "concurrent_ranks": world}


# ------------------------------------------------------------------------------------------------------------------------
# the full path (config 3)
# ------------------------------------------------------------------------------------------------------------------------
def full_path_weights():
    from oryon_b200 import synth_backbone as sb
    sd = sb.oryon_state_dict(11)
    sd["decoder.head.bias"] = sd["decoder.head.bias"] + FULL["head_bias_shift"]
    return sd


def prompt_lists(n=FULL["distinct_prompts"]):
    """34 distinct prompt lists of 81 strings (object name + 80 templated prompts, datasets.py:515-532)."""
    return [[f"object{i}"] + [f"a photo number {t} of a object{i}." for t in range(80)] for i in range(n)]


def build_full_path(local, precision, mask_mode, pipelined=True, pairs_per_pass=32):
    from oryon_b200 import synth
    from oryon_b200.models.tokenizer import SimpleTokenizer
    from oryon_b200.net import Oryon
    from oryon_b200.pipeline import FPM_Pipeline
    from oryon_b200.utils.pointdsc.init import PointDSCSolver
    cfg = synth.POINTDSC_DEFAULT_CFG
    dev = f"cuda:{local}"
    model = Oryon(None, dev, state_dict=full_path_weights(), precision=precision, tokenizer=SimpleTokenizer(BPE_SYNTH),
                  max_pairs_per_pass=pairs_per_pass)
    solver = PointDSCSolver(synth.pointdsc_state_dict(300), in_dim=cfg["in_dim"], num_layers=cfg["num_layers"],
                            num_channels=cfg["num_channels"], num_iterations=cfg["num_iterations"], ratio=cfg["ratio"],
                            sigma_d=cfg["sigma_d"], k=cfg["k"], nms_radius=cfg["inlier_threshold"], device=dev)
    args = dict(device=dev, corrs_device="cpu", seed=1, dataset=dict(img_size=[224, 224], max_corrs=500),
                model=dict(image_encoder=dict(img_size=[192, 192])),
                test=dict(mask=mask_mode, src_sampling=5000, solver="pointdsc", n_corrs=500, dist_th=0.25, mask_threshold=0.5,
                          pipelined=pipelined))
    return FPM_Pipeline(args, test_model=True, model=model, pointdsc_solver=solver), model


def host_batches(n_batches, B):
    """`n_batches` distinct synthetic batches in the schema of the reference's collate (datasets.py:202-245), as the B200
    collate stages them on the host: pinned float RGB, pinned uint8 masks, the depth frames stacked into one pinned tensor."""
    from oryon_b200 import synth
    out = []
    for k in range(n_batches):
        b = synth.synthetic_batch(100 + k, B)
        b.pop("prompt_tokens")
        for key in ("anchor", "query"):
            b[key]["rgb"] = b[key]["rgb"].pin_memory()
            b[key]["mask"] = b[key]["mask"].pin_memory()
            b[key]["orig_depth"] = torch.stack(b[key]["orig_depth"]).pin_memory()
        out.append(b)
    return out


def to_device_batch(b, dev):
    out = {k: v for k, v in b.items() if k not in ("anchor", "query")}
    for key in ("anchor", "query"):
        out[key] = {k: (v.to(dev) if k in ("rgb", "mask", "orig_depth") else v) for k, v in b[key].items()}
    return out


def step_batch(batches, prompts, step_index, first_pair, B):
    """The batch of step `step_index`: a distinct synthetic batch (cycled) with the prompt lists of its pairs of the split."""
    b = dict(batches[step_index % len(batches)])
    idx = list(range(first_pair, first_pair + B))
    b["prompt"] = [prompts[p % len(prompts)] for p in idx]
    b["pair_index"] = idx
    return b


def e2e_bytes_per_step(batch, B, n_corrs=500, featmap=192):
    """Host<->device bytes of one test_step, counted from the tensors that are copied: inputs up (RGB, masks, depth frames,
    the drawn row table), per-pair results down (nearest-neighbour distances for the draws on the CPU generator, ROI / point
    counts, poses, IoUs)."""
    h2d = sum(batch[k][f].numel() * batch[k][f].element_size() for k in ("anchor", "query") for f in ("rgb", "mask", "orig_depth"))
    h2d += B * n_corrs * 4
    d2h = B * featmap * featmap * 4 + 2 * B * 4 + B * 4 + B * 16 * 4 + 2 * B * 4
    return int(h2d), int(d2h)


def cpu_full_path_pairs(n_pairs, threads, warm=True, mask_mode=FULL["mask"]):
    """The reference's whole test step, pair by pair, on the host cores through the oracle port (backbone_oracle.oryon_forward
    restates net.py:142-167 -- including the prompt encoding the reference repeats on every step, net.py:147 --
    oryon_oracle.post_network_step restates pipeline.py:311-355).  Same synthetic weights / inputs / mask mode as the GPU arm.
    Returns the per-pair wall times (after one untimed warm-up forward on a single pair when `warm`)."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import backbone_oracle as bo
    import oryon_oracle as oracle
    from oryon_b200 import synth
    torch.set_num_threads(threads)
    w = full_path_weights()
    batch = synth.synthetic_batch(100, max(1, min(n_pairs, 4)))
    swin = bo.guidance_backbone(w)
    pw = synth.pointdsc_state_dict(300)

    def one(i):
        sl = slice(i, i + 1)
        t0 = time.perf_counter()
        out = bo.oryon_forward(w, batch["anchor"]["rgb"][sl], batch["query"]["rgb"][sl], batch["prompt_tokens"][sl], swin=swin)
        t1 = time.perf_counter()
        sub = {k: ({f: v[sl] for f, v in batch[k].items()} if isinstance(batch[k], dict) else batch[k][sl]) for k in ("anchor", "query")}
        rows = oracle.post_network_step({k: out[k] for k in ("featmap_a", "featmap_q", "mask_a", "mask_q")}, sub, pw,
                                        synth.POINTDSC_DEFAULT_CFG, mask_mode=mask_mode)
        t2 = time.perf_counter()
        return t1 - t0, t2 - t1, rows[0]["status"]

    torch.manual_seed(1)
    if warm:
        one(0)
    return [one(i % batch["anchor"]["rgb"].shape[0]) for i in range(n_pairs)]


def run_reference(args):
    """--impl reference: the CPU port of the reference's full step, all host threads; one step = ONE pair of the workload."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    if args.warmup > 0:
        cpu_full_path_pairs(min(args.warmup, 1), threads, warm=False)     # one whole pair warms every code path up
    t0 = time.perf_counter()
    per = cpu_full_path_pairs(args.steps, threads, warm=False)
    dt = time.perf_counter() - t0
    pairs_per_s = args.steps / dt
    sample = (f"{args.steps} steps of ONE pair each (1 of the {FULL['B']} pairs of a batch), all {threads} host threads, float32 CPU PyTorch; "
              f"network {statistics.mean(p[0] for p in per):.2f} s + post-network {statistics.mean(p[1] for p in per):.2f} s per pair; "
              "the prompt set is re-encoded on every step as the reference does (net.py:147)")
    line = {"impl": "reference", "metric": "image-pairs/sec", "value": pairs_per_s, "unit": "pairs/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3 / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(),
            "note": "the reference is Python/PyTorch and cannot travel to the GPU box: this is the oracle port (pinned to the reference's own "
                    "outputs by tests/golden); a step of this arm is one pair, not a 32-pair batch: value is pairs/s either way",
            "status": {s: sum(p[2] == s for p in per) for s in ("ok", "no_corrs", "invalid_mask")},
            "cpu_baseline": {"value": pairs_per_s, "unit": "pairs/s", "cores": threads, "kind": "port", "sample": sample},
            "e2e": {"value": pairs_per_s, "unit": "pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    _emit(line)


# ------------------------------------------------------------------------------------------------------------------------
# matcher regions (configs 2 and 5) and the optional baselines
# ------------------------------------------------------------------------------------------------------------------------
def cpu_port_sample(rows, threads, D=C2["D"], n=C2["H"] * C2["W"], repeats=1, seed=0):
    """The oracle's matcher (reference utils/pcd.py:202-204 restated) on `rows` anchor rows of one pair against all `n`
    query positions.  Seconds per call (best of `repeats`)."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import oryon_oracle as oracle
    torch.set_num_threads(threads)
    g = torch.Generator().manual_seed(seed)
    f1 = torch.randn(rows, D, generator=g)
    f2 = torch.randn(n, D, generator=g)
    best = float("inf")
    for _ in range(repeats):
        t0 = time.perf_counter()
        oracle.match_rows(f1, f2, row_chunk=64)
        best = min(best, time.perf_counter() - t0)
    return best


def eager_torch_sample(device, dtype, rows, D=C2["D"], n=C2["H"] * C2["W"], row_chunk=256, repeats=3, seed=0):
    """`--eager-baseline`: the reference's own formulation of the matcher (utils/pcd.py:202-204: broadcast `cosine_similarity`,
    `amin`, `argmin`) as plain PyTorch eager ops on `device` -- what the unmodified reference does with `corrs_device='cuda'`
    (float16, :195-197) or with float32 features moved to the GPU -- in row chunks, because the N1 x N2 x D broadcast of one
    config-2 pair (94 GB in float16) does not fit.  Seconds per call, best of `repeats`."""
    g = torch.Generator().manual_seed(seed)
    f1 = torch.randn(rows, D, generator=g).to(device=device, dtype=dtype)
    f2 = torch.randn(n, D, generator=g).to(device=device, dtype=dtype)
    cos = torch.nn.functional.cosine_similarity
    sync = torch.cuda.synchronize if torch.device(device).type == "cuda" else (lambda *a: None)
    best = float("inf")
    for _ in range(repeats + 1):                   # the first pass warms the allocator up
        sync()
        t0 = time.perf_counter()
        mins, args_ = [], []
        for r0 in range(0, rows, row_chunk):
            d = 0.5 * (-1 * cos(f1[r0:r0 + row_chunk].unsqueeze(1), f2.unsqueeze(0), dim=2) + 1)
            mins.append(torch.amin(d, dim=1))
            args_.append(torch.argmin(d, dim=1))
        torch.cat(mins), torch.cat(args_)
        sync()
        best = min(best, time.perf_counter() - t0)
    return best


def eager_baseline(device, rows=None):
    n = C2["H"] * C2["W"]
    rows = rows or n
    out = {"what": "reference formulation (utils/pcd.py:202-204) as PyTorch eager ops on the GPU, row chunks of 256; not the product path",
           "sample": f"{rows} of {n} anchor rows of one pair x all {n} query positions, D={C2['D']}"}
    for name, dt in (("f32", torch.float32), ("f16", torch.float16)):
        sec = eager_torch_sample(device, dt, rows)
        out[name] = {"pairs_per_s": (rows / n) / sec, "ms_per_pair": sec * 1e3 * n / rows}
    return out


def measured_traffic(kernel, cfg_name):
    """`roofline.traffic`: dram__bytes_read.sum + dram__bytes_write.sum per launch from the `ncu --set full` capture that
    tools/ncu_traffic.py regenerates with THIS bench command; used only while the capture's source fingerprint equals the
    library the bench is running (a stale capture reads as null, never as a number)."""
    try:
        from oryon_b200 import build as _build
        rec = json.load(open(TRAFFIC_FILE))
        if rec.get("fingerprint") != _build._fingerprint():
            return None, f"{os.path.relpath(TRAFFIC_FILE, ROOT)} was captured from other sources than the ones running: null"
        ent = rec["kernels"][f"{kernel}:{cfg_name}"]
        return float(ent["dram_bytes_read"] + ent["dram_bytes_write"]), (f"ncu --set full capture of this bench command "
                                                                         f"({os.path.relpath(TRAFFIC_FILE, ROOT)}, {rec.get('when')})")
    except Exception as e:
        return None, f"no ncu capture for this build ({type(e).__name__}): null"


def matcher_region(cfg, cfg_name, local, world, rank, barrier, peaks, min_seconds, e2e_steps=0):
    """Dense matcher (oryon_match_nn) on `cfg`, inputs resident: repeats the batch for >= `min_seconds`, per-kernel CUDA events
    from the library (`oryon_profile_*`, recorded on the launching stream) -> pairs/s and the roofline entry of match_tc."""
    import torch.distributed as dist
    from oryon_b200 import _lib, synth
    from oryon_b200.utils import pcd
    dev = torch.device("cuda", local)
    B, D, H, W = cfg["B"], cfg["D"], cfg["H"], cfg["W"]
    n = H * W
    fa, fq, perm = synth.permuted_feature_batch(1000 + rank, B, D, H, W, noise=0.1, device=f"cuda:{local}")
    idx = torch.empty(B, n, dtype=torch.int32, device=dev)
    dst = torch.empty(B, n, dtype=torch.float32, device=dev)

    def step():
        pcd.match_nn(fa, fq, out=(idx, dst))

    for _ in range(3):
        step()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        step()
    e1.record()
    torch.cuda.synchronize()
    steps = max(5, int(min_seconds / (e0.elapsed_time(e1) * 1e-3 / 3)) + 1)
    t = torch.tensor([float(steps)], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    steps = int(t.item())
    launches_per_step = pcd.match_last_stats(dev)["kernels_launched"]
    _lib.profile_enable(local, True)
    _lib.profile_read(local)
    with ClockSampler(local) as clocks:
        barrier()
        e0.record()
        for _ in range(steps):
            step()
        e1.record()
        barrier()
    ms = e0.elapsed_time(e1)
    prof = _lib.profile_read(local)
    _lib.profile_enable(local, False)
    t = torch.tensor([ms], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_max = float(t.item())
    ok = bool(torch.equal(idx.long(), perm))      # the planted permutation is recovered: the timed work is the real work
    stats = pcd.match_last_stats(dev)
    flops, bytes_ = algorithmic_work(B, D, n, n)
    tc_ms, tc_n = prof.get("match_tc", (0.0, 0))
    tc_per_step = max(tc_n // max(steps, 1), 1)
    tc_avg = tc_ms / max(tc_n, 1)
    achieved_tf = flops / (tc_avg * tc_per_step * 1e-3) / 1e12 if tc_avg > 0 else None
    region_s = ms_max * 1e-3
    burst, sustained = peaks.get("bf16_tflops") or 1650.0, peaks.get("bf16_tflops_sustained") or 1400.0
    use_sustained = region_s >= 1.0
    peak_tf = sustained if use_sustained else burst
    hbm_peak = peaks.get("hbm_gbs") or 6650.0
    traffic, traffic_src = measured_traffic("match_tc_kernel", cfg_name)      # the tensor-core pass (match_tc2_kernel: CTA pairs)
    step_ms = ms_max / steps
    out = {
        "workload": f"{cfg_name}: B={B} pairs/GPU, D={D}, {H}x{W} ({n} positions/image), dense all-pairs NN matching ({4 * H}x{4 * W} frames at "
                    f"stride 4); inputs {2 * B * D * n * 4 / 1e6:.0f} MB per step (> 126 MB L2)",
        "value": world * B * steps / region_s, "unit": "pairs/s", "steps": steps, "region_s": region_s, "ms_per_step": step_ms,
        "results_ok": ok, "clocks": clocks.summary(), "kernels_ms_per_step": {str(k): v[0] / steps for k, v in prof.items()},
        "match_stats_last_step": stats, "gpu_launches_per_step": int(launches_per_step),
        "roofline": {"kernel": "match_tc_kernel", "config": cfg_name, "bound": "tensor", "achieved": achieved_tf, "peak": peak_tf, "unit": "TFLOP/s",
                     "frac": (achieved_tf / peak_tf if achieved_tf else None), "traffic": traffic, "traffic_source": traffic_src,
                     "peak_source": ("MEASURED_PEAKS.json " if peaks else "fallback (B200_PROFILING.md) ") +
                                    (f"bf16_tflops_sustained: the kernel was timed inside a {region_s:.1f} s region of back-to-back steps"
                                     if use_sustained else f"bf16_tflops (burst): the timed region is only {region_s * 1e3:.0f} ms"),
                     "frac_of_burst_peak": (achieved_tf / burst if achieved_tf else None),
                     "frac_of_sustained_peak": (achieved_tf / sustained if achieved_tf else None),
                     "algorithmic_flops_per_launch": flops / tc_per_step, "launch_ms": tc_avg, "launches_per_step": tc_per_step,
                     "launches_timed": int(tc_n), "kernel_share_of_step": (tc_avg * tc_per_step / step_ms if step_ms else None),
                     "hbm": {"algorithmic_bytes_per_step": bytes_, "achieved_gbs": bytes_ / (step_ms * 1e-3) / 1e9,
                             "peak_gbs": hbm_peak, "frac": bytes_ / (step_ms * 1e-3) / 1e9 / hbm_peak}},
    }
    if e2e_steps:
        ha, hq = fa.cpu().pin_memory(), fq.cpu().pin_memory()
        h_idx = torch.empty(B, n, dtype=torch.int32).pin_memory()
        h_dst = torch.empty(B, n, dtype=torch.float32).pin_memory()

        def e2e_step():
            i, d = pcd.match_nn_streamed(ha, hq, chunk_pairs=4)   # H2D of both maps inside, overlapped with the kernels
            h_idx.copy_(i, non_blocking=True)
            h_dst.copy_(d, non_blocking=True)
            torch.cuda.synchronize()

        for _ in range(2):
            e2e_step()
        barrier()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            e2e_step()
        barrier()
        te = torch.tensor([time.perf_counter() - t0], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(te, op=dist.ReduceOp.MAX)
        out["e2e"] = {"value": world * B * e2e_steps / float(te.item()), "unit": "pairs/s",
                      "h2d_bytes_per_step": int(ha.numel() * 4 + hq.numel() * 4), "d2h_bytes_per_step": int(h_idx.numel() * 8),
                      "steps": e2e_steps, "ok": bool(torch.equal(h_idx.long(), perm.cpu()))}
        del ha, hq
    del fa, fq, idx, dst
    torch.cuda.empty_cache()
    return out


_REAL_STDOUT = None


def _claim_stdout():
    """The contract is ONE JSON line on stdout.  Libraries print there too (NCCL announces its version from C code when a
    communicator is created), so file descriptor 1 is pointed at stderr for the whole run and the line is written to the
    saved descriptor at the end."""
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)


def _emit(line: dict):
    data = (json.dumps(line) + "\n").encode()
    sys.stdout.flush()
    if _REAL_STDOUT is None:
        os.write(1, data)
    else:
        os.write(_REAL_STDOUT, data)


def main():
    _claim_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=8)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--precision", type=int, default=2, choices=[1, 2, 3],
                    help="network GEMM precision (include/oryon_b200.h): 3 = three fp16 products everywhere; 2 (default, parity-tested to the same "
                         "1e-3 gate) = the CLIP vision linear layers with fp8 cross terms, three products elsewhere; 1 is NOT a valid bench mode")
    ap.add_argument("--mask", default=FULL["mask"], choices=["predicted", "oracle"])
    ap.add_argument("--cpu-pairs", type=int, default=3, help="pairs of the cpu_baseline sample (N = 1 only)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-matcher", action="store_true", help="skip the config-2 / config-5 matcher regions (roofline = null)")
    ap.add_argument("--matcher-only", action="store_true", help="only one matcher region (config 2; used by tools/ncu_traffic.py)")
    ap.add_argument("--matcher-config", type=int, default=2, choices=[2, 5], help="with --matcher-only: BASELINE config 2 or 5")
    ap.add_argument("--matcher-seconds", type=float, default=2.0, help="minimum length of the config-2 roofline region")
    ap.add_argument("--eager-baseline", action="store_true", help="also time the reference's matcher formulation as PyTorch eager ops on the GPU")
    ap.add_argument("--pairs-per-pass", type=int, default=32, help="pairs per network pass inside a step (activation arena size; 16 = round 1's "
                                                                     "setting: twice the launches, 2 %% slower, profiles/r02_bench_run_h_*.json)")
    ap.add_argument("--no-pipeline", action="store_true", help="A/B: run each batch's post-network tail before the next network pass instead of under it")
    ap.add_argument("--no-affinity", action="store_true", help="do not pin the rank to its GPU's CPU set (A/B for the e2e scaling)")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    args.warmup = max(args.warmup, 3)

    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product path has no CPU fallback")
    affinity = {"applied": False, "note": "--no-affinity"} if args.no_affinity else bind_rank_to_gpu_locality(local, max(world, 1))
    if "OMP_NUM_THREADS" not in os.environ:
        torch.set_num_threads(2)       # the loop's own CPU ops are the per-pair multinomial draws: tiny
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    from oryon_b200 import _lib, sharding
    from oryon_b200.net import gemm_counters

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass

    if args.matcher_only:
        m2 = matcher_region(C5 if args.matcher_config == 5 else C2, f"config{args.matcher_config}", local, world, rank, barrier, peaks,
                            args.matcher_seconds)
        if rank == 0:
            _emit({"matcher_only": True, **m2})
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- the full path -----------------------------------------------------------------------------------------------
    B, K, W = FULL["B"], args.steps, args.warmup
    pipe, model = build_full_path(local, args.precision, args.mask, pipelined=not args.no_pipeline, pairs_per_pass=args.pairs_per_pass)
    prompts = prompt_lists()
    hb = host_batches(FULL["distinct_batches"], B)
    db = [to_device_batch(b, dev) for b in hb]
    model.encode_prompt(prompts)                      # the 34 prompt sets of the split: encoded once, cached (Oryon.encode_prompt)
    # this rank's share of the timed pairs: K steps of B pairs per rank (weak scaling), contiguous in the split's order
    total_pairs = world * K * B
    mine = sharding.shard_pairs(total_pairs, rank, world)
    assert len(mine) == K * B

    def run_steps(batches, n_steps, first_step=0, collect=None):
        """`n_steps` test steps and the flush of the last one (pipelined mode: a step returns the rows of the batch before it)."""
        done = 0
        for s in range(n_steps + 1):
            if s < n_steps:
                p0 = (mine.start + (first_step + s) * B) % FULL["split_pairs"]
                rows = pipe.test_step(step_batch(batches, prompts, first_step + s, p0, B), first_step + s)
            else:
                rows = pipe.flush()
            if rows and collect is not None:
                collect.append(sharding.encode_rows(list(range(mine.start + done * B, mine.start + (done + 1) * B)), rows))
            done += 1 if rows else 0
        assert done == n_steps, (done, n_steps)

    pipe.on_test_start()
    run_steps(db, W)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    pipe.on_test_start()
    with ClockSampler(local) as clocks:
        barrier()
        e0.record()
        run_steps(db, K)
        e1.record()
        barrier()
    t = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_max = float(t.item())
    status_resident = {s: sum(r["status"] == s for r in pipe.rows) for s in ("ok", "no_corrs", "invalid_mask")}

    # ---- end to end: host batches in, gathered result rows out ----------------------------------------------------------
    run_steps(hb, 2)
    pipe.on_test_start()
    local_rows = []
    barrier()
    t0 = time.perf_counter()
    run_steps(hb, K, collect=local_rows)
    table = sharding.gather_rows(torch.cat(local_rows).to(dev), total_pairs).cpu()     # the path's only collective
    barrier()
    te = torch.tensor([time.perf_counter() - t0], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_s = float(te.item())
    records = sharding.decode_rows(table)
    status_e2e = {s: sum(r["status"] == s for r in records) for s in sharding.STATUS}
    h2d_b, d2h_b = e2e_bytes_per_step(hb[0], B)
    e2e = {"value": total_pairs / e2e_s, "unit": "pairs/s", "h2d_bytes_per_step": h2d_b, "d2h_bytes_per_step": d2h_b, "steps": K,
           "seconds": e2e_s, "pairs_gathered": len(records), "rows_in_pair_order": [r["pair_index"] for r in records] == list(range(total_pairs)),
           "status": status_e2e,
           "what": "FPM_Pipeline.test_step on HOST batches (pinned RGB / masks / depth frames, prompt strings -> tokenizer + cached text "
                   "tower), pose rows back on the host, then ONE all_gather of the 16-float result rows (sharding.gather_rows)"}

    # ---- one profiled step: per-kernel split, launches, GEMM rate --------------------------------------------------------
    _lib.profile_enable(local, True)
    _lib.profile_read(local)
    gemm_counters(local)
    torch.cuda.synchronize()
    t1 = time.perf_counter()
    run_steps(db, 1)
    torch.cuda.synchronize()
    profiled_ms = (time.perf_counter() - t1) * 1e3
    prof = _lib.profile_read(local)
    _lib.profile_enable(local, False)
    n_gemm, gemm_flops, gemm_tensor_flops = gemm_counters(local, with_tensor_flops=True)
    launches_per_step = int(sum(v[1] for v in prof.values()))
    probe = h2d_probe(dev, world, barrier)

    # ---- matcher regions: config 2 (the roofline kernel's configuration) and config 5 ------------------------------------
    m2 = m5 = None
    if not args.no_matcher:
        del db
        torch.cuda.empty_cache()
        m2 = matcher_region(C2, "config2", local, world, rank, barrier, peaks, args.matcher_seconds, e2e_steps=8)
        m5 = matcher_region(C5, "config5", local, world, rank, barrier, peaks, 1.0)

    if rank == 0:
        passes = gemm_tensor_flops / gemm_flops if gemm_flops else 0.0   # tensor-pipe products per algorithmic product, FLOP-weighted
        gemm_ms = prof.get("gemm_tc", (0.0, 0))[0]
        gemm_tf = gemm_flops / (gemm_ms * 1e-3) / 1e12 if gemm_ms else None
        sustained = peaks.get("bf16_tflops_sustained") or 1400.0
        line = {
            "metric": "image-pairs/sec", "value": total_pairs / (ms_max * 1e-3), "unit": "pairs/s", "n_gpus": world,
            "steps": K, "warmup": W, "ms_per_step": ms_max / K, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None,
            "dtype": DTYPE_NOTE[args.precision],
            "data": "synthetic", "config": workload_config(),
            "timed_region_s": ms_max * 1e-3, "status": status_resident,
            "clocks": clocks.summary(),
            "e2e": e2e,
            "gpu_launches": launches_per_step * K,
            "gpu_launches_per_step": launches_per_step,
            "kernels_ms_per_step": {str(k): round(v[0], 3) for k, v in prof.items()},
            "profiled_step_ms": profiled_ms,
            "network_gemm": {"launches_per_step": n_gemm, "algorithmic_tflop_per_step": gemm_flops / 1e12, "ms_per_step": gemm_ms,
                             "algorithmic_tflops": gemm_tf, "tensor_pipe_tflops": (passes * gemm_tf if gemm_tf else None),
                             "peak_tflops": sustained, "frac_algorithmic": (gemm_tf / sustained if gemm_tf else None),
                             "frac_tensor_pipe": (passes * gemm_tf / sustained if gemm_tf else None),
                             "tensor_products_per_product": passes, "note": GEMM_NOTE[args.precision]},
            "affinity": affinity, "h2d_probe": probe,
            "roofline": (m2["roofline"] if m2 else None),
            "roofline_config5": (m5["roofline"] if m5 else None),
            "matcher_config2": ({k: v for k, v in m2.items() if k != "roofline"} if m2 else None),
            "matcher_config5": ({k: v for k, v in m5.items() if k != "roofline"} if m5 else None),
        }
        if not args.no_cpu_baseline and world == 1:
            threads = os.cpu_count() or 1
            try:
                per = cpu_full_path_pairs(args.cpu_pairs, threads, warm=True)
                sec = sum(p[0] + p[1] for p in per)
                line["cpu_baseline"] = {"value": len(per) / sec, "unit": "pairs/s", "cores": threads, "kind": "port",
                                        "sample": f"{len(per)} pairs of the same workload after one warm-up pair, {sec:.1f} s measured: network "
                                                  f"{statistics.mean(p[0] for p in per):.2f} s + post-network {statistics.mean(p[1] for p in per):.2f} s per pair "
                                                  "(oracle port of the reference's step, float32 CPU PyTorch)",
                                        "status": [p[2] for p in per]}
            except Exception as e:
                line["cpu_baseline"] = {"error": repr(e)}
            if m2 is not None:
                n = C2["H"] * C2["W"]
                cpu_port_sample(32, threads)
                sec = cpu_port_sample(2048, threads)
                line["matcher_config2"]["cpu_baseline"] = {"value": (2048 / n) / sec, "unit": "pairs/s", "cores": threads, "kind": "port",
                                                           "sample": f"2048 of {n} anchor rows of one pair x all {n} query positions; {sec:.1f} s measured, scaled by {n / 2048:.1f}"}
        if args.eager_baseline and world == 1:
            try:
                line["gpu_eager_baseline"] = eager_baseline(f"cuda:{local}")
            except Exception as e:
                line["gpu_eager_baseline"] = {"error": repr(e)}
        _emit(line)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
