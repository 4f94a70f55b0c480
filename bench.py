#!/usr/bin/env python
"""bench.py — headline benchmark of the Online-RepNeRV frame-fitting hot path on B200.

Metric (BASELINE.json): training frames/s of the ERB model on the Bunny-shaped 720p configuration
(configs[1]: ERB, 132 x 720 x 1280 synthetic frames, fc_hw_dim 9_16_26, strides 5 2 2 2 2, batch 1 per GPU).
One "step" = one pass of the hot path over one frame per GPU: PE + stem, 5 x (ERB fold + conv3x3 +
PixelShuffle + SiLU), RGB head, Fusion6 loss + backward, full backward, (gradient all-reduce), fused Adam,
PSNR + MS-SSIM — i.e. one iteration of reference main_train.py:229-254.

  python bench.py [--gpus N] [--steps K] [--warmup W]            our arm   (torchrun for N > 1)
  python bench.py --impl reference [--steps K] [--warmup W]      CPU arm   (oracle port of the reference)

Prints ONE JSON line (rank 0).  `value` is timed with inputs resident in HBM; `e2e` is the same step driven
through the public API with pinned HOST frames: H2D of the uint8 frame + index and D2H of the metrics every
step, inside the timed region.
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

WORKLOAD = dict(name="ERB Bunny-shaped synthetic 132x720x1280 (BASELINE configs[1])", n_frames=132, H=720, W=1280,
                embed='1.25_40', stem_dim_num='512_1', fc_hw_dim='9_16_26', expansion=1, reduction=2,
                lower_width=96, strides=[5, 2, 2, 2, 2], branch_type='ERB', loss_type='Fusion6', lr=5e-4,
                epochs=300, warmup_ratio=0.2, beta=0.5)
# algorithmic work: 605.7 GFLOP per training step of one 720p frame (+4.3 GFLOP fold)
STEP_GFLOP = 605.7
L4_GEMM_GFLOP = 2.0 * (360 * 640) * 384 * 864 / 1e9     # one conv pass of block 4 (fprop = dgrad = wgrad)


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            d = json.load(f)
        return dict(bf16=d.get("bf16_tflops", 1590.0), bf16_sustained=d.get("bf16_tflops_sustained", 1400.0),
                    hbm=d.get("hbm_gbs", 6650.0), source="measured")
    return dict(bf16=1590.0, bf16_sustained=1400.0, hbm=6650.0, source="fallback")


class ClockSampler:
    """SM clock / throttle reasons sampled DURING the timed region, in-process through NVML (nvidia_ml_py).
    (An external `nvidia-smi -lms 100` poller next to the persistent tcgen05 kernels coincided with GPU-side hangs
    on this driver — see DESIGN.md section 6 — so nothing is spawned here.)"""

    def __init__(self, gpu_index, period_s=0.02):
        self.gpu, self.period, self.rows, self.thread = gpu_index, period_s, [], None
        self._stop = threading.Event()
        self.mode = os.environ.get("ONR_BENCH_SAMPLER", "nvml")      # nvml | smi | off

    def _nvml_loop(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            h = pynvml.nvmlDeviceGetHandleByIndex(self.gpu)
            mx = pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM)
            bits = {"hw_slowdown": pynvml.nvmlClocksThrottleReasonHwSlowdown,
                    "hw_thermal_slowdown": pynvml.nvmlClocksThrottleReasonHwThermalSlowdown,
                    "sw_thermal_slowdown": pynvml.nvmlClocksThrottleReasonSwThermalSlowdown,
                    "sw_power_cap": pynvml.nvmlClocksThrottleReasonSwPowerCap}
            while not self._stop.is_set():
                sm = pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)
                r = pynvml.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                self.rows.append((float(sm), float(mx), [n for n, bit in bits.items() if r & bit]))
                self._stop.wait(self.period)
        except Exception as exc:                                       # noqa: BLE001
            self.rows.append((None, None, [f"nvml unavailable: {exc}"]))

    def _smi_loop(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        while not self._stop.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--id={self.gpu}", f"--query-gpu={q}",
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
                c = [x.strip() for x in out.strip().split(",")]
                self.rows.append((float(c[0]), float(c[1]), [n for n, v in zip(names, c[2:6]) if v.lower() == "active"]))
            except Exception as exc:                                   # noqa: BLE001
                self.rows.append((None, None, [f"nvidia-smi unavailable: {exc}"]))
            self._stop.wait(max(self.period, 0.5))

    def start(self):
        if self.mode == "off":
            return
        self.thread = threading.Thread(target=self._smi_loop if self.mode == "smi" else self._nvml_loop, daemon=True)
        self.thread.start()

    def stop(self):
        if self.thread is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["sampler off"], "samples": 0}
        self._stop.set()
        self.thread.join(timeout=3)
        sm = sorted(r[0] for r in self.rows if r[0] is not None)
        mx = [r[1] for r in self.rows if r[1] is not None]
        reasons = sorted({x for r in self.rows for x in r[2]})
        busy = sm[len(sm) // 2:]           # idle samples before the first kernel pull the median down
        return {"sm_mhz": statistics.median(busy) if busy else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm), "source": self.mode}


def make_args(world):
    w = WORKLOAD
    return argparse.Namespace(loss_type=w['loss_type'], lr=w['lr'], lr_type='cosine', epochs=w['epochs'],
                              warmup=int(w['warmup_ratio'] * w['epochs']), beta=w['beta'], batchSize=1)


# ----------------------------------------------------------------------------------------------- CPU arm
def cpu_steps(n_steps, warmup, threads=None):
    """Times `n_steps` training steps of ONE 720p frame with the oracle port of the reference on host cores."""
    import torch
    from oracle import nerv_oracle as O
    from orepnerv.data import synthetic_clip
    if threads:
        torch.set_num_threads(threads)
    w = WORKLOAD
    fh, fw, fd = [int(x) for x in w['fc_hw_dim'].split('_')]
    cfg = dict(fc_h=fh, fc_w=fw, fc_dim=fd, strides=w['strides'], sigmoid=False)
    sd = reference_shaped_state(w)
    frames = synthetic_clip(2, w['H'], w['W']).float().div(255)
    state = {}
    times = []
    for i in range(warmup + n_steps):
        pos = torch.tensor([(i % 2) / w['n_frames']])
        embed = O.pos_encoding(pos, 1.25, 40)
        t0 = time.perf_counter()
        lr = O.lr_at(0, i, w['n_frames'], w['lr'], int(w['warmup_ratio'] * w['epochs']), w['epochs'])
        sd, state, loss, img, _ = O.train_step(sd, state, embed, frames[i % 2:i % 2 + 1], cfg, lr, i + 1)
        _ = O.psnr(img, frames[i % 2:i % 2 + 1])
        _ = O.ms_ssim(img, frames[i % 2:i % 2 + 1])
        dt = time.perf_counter() - t0
        if i >= warmup:
            times.append(dt)
    return sum(times), torch.get_num_threads()


def reference_shaped_state(w):
    """Random-init ERB state dict with the reference's parameter shapes (CPU, no GPU needed)."""
    import torch
    from orepnerv.model import Generator
    torch.manual_seed(1)
    gen = Generator(embed_length=80, stem_dim_num=w['stem_dim_num'], fc_hw_dim=w['fc_hw_dim'], expansion=w['expansion'],
                    num_blocks=1, norm='none', act='swish', bias=True, reduction=w['reduction'], conv_type='conv',
                    stride_list=w['strides'], sin_res=True, lower_width=w['lower_width'], sigmoid=False,
                    deploy=False, branch_type=w['branch_type'])
    return {k: v.detach().clone() for k, v in gen.state_dict().items()}


def run_reference(opts):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    total, cores = cpu_steps(opts.steps, opts.warmup)
    fps = opts.steps / total
    line = {
        "impl": "reference", "metric": "train frames/s (720p ERB)", "value": fps, "unit": "frames/s",
        "n_gpus": opts.gpus, "steps": opts.steps, "warmup": opts.warmup, "ms_per_step": 1000.0 * total / opts.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD['name'], "batch_per_gpu": 1, "note": "CPU arm: oracle port of the "
                   "reference step (reference is Python/PyTorch and /root/reference does not travel to the GPU box)"},
        "cpu_baseline": {"value": fps, "unit": "frames/s", "cores": cores, "kind": "port",
                         "sample": f"{opts.steps} ERB training steps of one 720x1280 frame (fwd, Fusion6, bwd, Adam, PSNR, MS-SSIM)"},
        "e2e": {"value": fps, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------------------------- GPU arm
def _mark(msg):
    if os.environ.get("ONR_BENCH_VERBOSE"):
        print(f"[bench rank {os.environ.get('RANK', '0')}] {msg}", file=sys.stderr, flush=True)


def run_ours(opts):
    import torch
    import torch.distributed as dist
    from orepnerv import _lib, sharding
    from orepnerv.data import synthetic_clip
    from orepnerv.model import Generator
    from orepnerv.trainer import FrameFitter
    from orepnerv.utils import PositionalEncoding

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != opts.gpus:
        if world == 1 and opts.gpus > 1:
            raise SystemExit("launch with torchrun --nproc-per-node N for --gpus N > 1")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    lib = _lib.lib()
    _mark("process group + library ready")
    w = WORKLOAD
    args = make_args(world)
    torch.manual_seed(1)
    pe = PositionalEncoding(w['embed'])
    gen = Generator(embed_length=pe.embed_length, stem_dim_num=w['stem_dim_num'], fc_hw_dim=w['fc_hw_dim'],
                    expansion=w['expansion'], num_blocks=1, norm='none', act='swish', bias=True,
                    reduction=w['reduction'], conv_type='conv', stride_list=w['strides'], sin_res=True,
                    lower_width=w['lower_width'], sigmoid=False, deploy=False, branch_type=w['branch_type']).to(dev)
    n_frames = opts.frames or w['n_frames']
    clip = synthetic_clip(n_frames, w['H'], w['W'], device=dev)                 # uint8, resident in HBM
    spe = sharding.steps_per_epoch(n_frames, world)
    fit = FrameFitter(gen, pe, args, world_size=world, data_size=n_frames, steps_per_epoch=spe,
                      use_graph=not opts.no_graph)
    t_all = torch.arange(n_frames, dtype=torch.float32, device=dev) / n_frames
    order = []
    ep = 0
    while len(order) < 2 * (opts.steps + opts.warmup) + 8:
        order += sharding.shard_indices(n_frames, world, rank, ep)
        ep += 1

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    _mark("model, clip and fitter built")
    # ---------------- device-resident timing (`value`) ----------------
    it = iter(order)
    for _ in range(opts.warmup):
        i = next(it)
        fit.step(clip[i:i + 1], t_all[i:i + 1])
    _mark("warm-up done")
    barrier()
    launches0 = lib.onr_launch_count()
    sampler = ClockSampler(local_rank)
    sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(opts.steps):
        i = next(it)
        fit.step(clip[i:i + 1], t_all[i:i + 1])
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    clocks = sampler.stop()
    launches_host = lib.onr_launch_count() - launches0
    out_last = fit.out.clone()
    if world > 1:
        tms = torch.tensor([ms], device=dev)
        dist.all_reduce(tms, op=dist.ReduceOp.MAX)
        ms = tms.item()
    value = world * opts.steps / (ms / 1000.0)

    _mark(f"timed region done: {ms:.2f} ms")
    # ---------------- end-to-end through the public API with host buffers (`e2e`) ----------------
    pinned_frames = [clip[i:i + 1].cpu().pin_memory() for i in order[:8]]
    pinned_t = [(torch.tensor([i], dtype=torch.float32) / n_frames).pin_memory() for i in order[:8]]
    host_out = torch.zeros(8, dtype=torch.float32).pin_memory()
    # public API for host-fed training: FrameFitter.host_pipeline_begin / step_host / host_pipeline_end.  Every step
    # uploads its own pinned uint8 frame + index (H2D) and downloads its own metrics (D2H) inside the timed region;
    # the copies run on a copy stream one step ahead / behind so they overlap the neighbouring steps' kernels.
    fit.host_pipeline_begin(pinned_frames[0], pinned_t[0])
    for k in range(3):
        fit.step_host(pinned_frames[(k + 1) % 8], pinned_t[(k + 1) % 8])
    fit.host_pipeline_end()
    barrier()
    f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    results = []
    f0.record()
    fit.host_pipeline_begin(pinned_frames[0], pinned_t[0])                   # H2D of step 0's inputs
    for k in range(opts.steps):
        nxt = (k + 1) % 8
        last = k == opts.steps - 1
        prev = fit.step_host(None if last else pinned_frames[nxt], None if last else pinned_t[nxt])
        if prev is not None:
            results.append(prev)                                              # D2H result of step k-1, on the host
    results.append(fit.host_pipeline_end())                                   # D2H result of the last step
    f1.record()
    barrier()
    assert len(results) == opts.steps
    host_out.copy_(results[-1])
    ms_e2e = f0.elapsed_time(f1)
    if world > 1:
        tms = torch.tensor([ms_e2e], device=dev)
        dist.all_reduce(tms, op=dist.ReduceOp.MAX)
        ms_e2e = tms.item()
    e2e = world * opts.steps / (ms_e2e / 1000.0)

    _mark("e2e done")
    # ---------------- kernels per step (claimed gpu_launches) ----------------
    c0 = lib.onr_launch_count()
    fit._body()
    torch.cuda.synchronize()
    per_step = lib.onr_launch_count() - c0

    # ---------------- roofline of the dominant kernels (block 4 convolution passes) ----------------
    peaks = measured_peaks()
    ex = fit.ex
    st = torch.cuda.current_stream().cuda_stream
    kern = {}
    for name, fn in (("conv_igemm_kernel<fprop,block4>", lambda: lib.onr_conv_plan_run(ex.fprop[4].handle, st)),
                     ("conv_igemm_kernel<dgrad,block4>", lambda: lib.onr_conv_plan_run(ex.dgrad[4].handle, st)),
                     ("wgrad_igemm_kernel<block4>", lambda: lib.onr_wgrad_plan_run(ex.wgrad[4].handle, st))):
        for _ in range(3):
            fn()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(10):
            fn()
        b.record()
        torch.cuda.synchronize()
        kern[name] = a.elapsed_time(b) / 10.0
    dom = max(kern, key=kern.get)
    achieved = L4_GEMM_GFLOP / kern[dom]            # GFLOP / ms = TFLOP/s
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "r01_dram_traffic.json")
    if os.path.exists(tpath):
        with open(tpath) as f:
            traffic = json.load(f).get(dom)
    roofline = {"bound": "tensor", "kernel": dom, "achieved": achieved, "peak": peaks['bf16'], "unit": "TFLOP/s",
                "frac": achieved / peaks['bf16'], "traffic": traffic, "traffic_unit": "bytes/launch (ncu dram read+write)",
                "peak_source": peaks['source'] + " bf16 burst",
                "algorithmic_gflop_per_launch": L4_GEMM_GFLOP,
                "launch_ms": {k: round(v, 4) for k, v in kern.items()},
                "step_tflops": STEP_GFLOP / (ms / opts.steps), "step_frac_of_sustained": STEP_GFLOP / (ms / opts.steps) / peaks['bf16_sustained']}

    # ---------------- reparameterised decode (BASELINE metric's second figure; reference main_eval.py decode loop) ----
    decode = None
    if rank == 0:
        import copy
        dep = copy.deepcopy(gen)
        for blk in dep.layers:
            blk.switch_to_deploy()                     # ERB branches folded into one 3x3 conv per block
        dep.eval()
        with torch.no_grad():
            embeds = [pe(t_all[i:i + 1]) for i in range(8)]
            for k in range(5):
                dep(embeds[k % 8])
            torch.cuda.synchronize()
            d0, d1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            n_dec = max(opts.steps, 20)
            d0.record()
            for k in range(n_dec):
                dep(embeds[k % 8])
            d1.record()
            torch.cuda.synchronize()
        ms_dec = d0.elapsed_time(d1) / n_dec
        decode = {"value": 1000.0 / ms_dec, "unit": "frames/s", "ms_per_frame": ms_dec, "batch": 1,
                  "what": "switch_to_deploy single-branch decode through Generator.__call__ (eager launches, packed "
                          "weights cached), device-timed; prune/quant only change weight values, not the kernels"}
        del dep
    _mark("decode done")

    if rank == 0:
        cpu_total, cores = cpu_steps(2, 1) if (world == 1 and not opts.no_cpu_baseline) else (None, None)
        line = {
            "metric": "train frames/s (720p ERB)", "value": value, "unit": "frames/s", "n_gpus": world,
            "steps": opts.steps, "warmup": opts.warmup, "ms_per_step": ms / opts.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": w['name'], "batch_per_gpu": 1, "global_batch": world,
                       "parallelism": f"frame-sharded dp{world}", "l2": "per-step working set (~0.9 GB of "
                       "activations) exceeds the 126 MB L2; no explicit flush", "cuda_graph": fit.graph is not None,
                       "metrics_every_step": "PSNR + MS-SSIM (reference main_train.py:253-254)"},
            "clocks": clocks,
            "e2e": {"value": e2e, "unit": "frames/s", "ms_per_step": ms_e2e / opts.steps,
                    "h2d_bytes_per_step": int(pinned_frames[0].numel() + 4), "d2h_bytes_per_step": 32},
            "gpu_launches": int(per_step * opts.steps), "kernels_per_step": int(per_step),
            "roofline": roofline,
            "last_step": {"loss": out_last[0].item(), "psnr": out_last[4].item(), "msssim": out_last[5].item()},
            "decode": decode,
        }
        if cpu_total is not None:
            line["cpu_baseline"] = {"value": 2 / cpu_total, "unit": "frames/s", "cores": cores, "kind": "port",
                                    "sample": "2 ERB training steps of one 720x1280 frame after 1 warm-up (oracle port)"}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def _arm_watchdog(seconds):
    """A wedged GPU wait must not hang the harness: dump every Python stack and exit non-zero."""
    import faulthandler
    faulthandler.enable()
    faulthandler.dump_traceback_later(seconds, exit=True)


def main():
    _arm_watchdog(int(os.environ.get("ONR_BENCH_WATCHDOG_S", "900")))
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--frames", type=int, default=0, help="clip length override (profiling runs only)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    opts = ap.parse_args()
    if opts.impl == "reference":
        run_reference(opts)
    else:
        opts.warmup = max(opts.warmup, 3)
        run_ours(opts)


if __name__ == "__main__":
    main()
