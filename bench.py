#!/usr/bin/env python
"""bench.py — headline benchmark of the Online-RepNeRV frame-fitting hot path on B200.

Metric (BASELINE.json): training frames/s of the ERB model on the Bunny-shaped 720p configuration
(configs[1]: ERB, 132 x 720 x 1280 synthetic frames, fc_hw_dim 9_16_26, strides 5 2 2 2 2, batch 1 per GPU).
One "step" = one pass of the hot path over one frame per GPU: PE + stem, 5 x (ERB fold + conv3x3 +
PixelShuffle + SiLU), RGB head, Fusion6 loss + backward, full backward, (gradient all-reduce), fused Adam,
PSNR + MS-SSIM — i.e. one iteration of reference main_train.py:229-254.

  python bench.py [--config c1|c2|c3|c4] [--gpus N] [--steps K] [--warmup W]   our arm (torchrun for N > 1)
  python bench.py --impl reference [--config ...] [--steps K] [--warmup W]     CPU arm (oracle port of the reference)
  python bench.py --impl torch-gpu [--config ...]    stock PyTorch (cuDNN/cuBLAS, TF32, cudnn.benchmark) on the
                                                     same B200: the GPU library baseline of SURVEY.md section 2

--config selects the BASELINE.json configuration (default c1 = configs[1], the one the metric is quoted on):
  c1  ERB S720  132 x 720 x 1280,  fc_hw_dim 9_16_26,  strides 5 2 2 2 2          (1 GPU headline)
  c2  ERB L720  132 x 720 x 1280,  fc_hw_dim 9_16_112 (NeRV-L width)              (2/4/8 GPUs)
  c3  ERB U1080 600 x 1080 x 1920, fc_hw_dim 9_16_26,  strides 5 3 2 2 2          (8 GPUs)
  c4  reparameterised single-branch decode of c1 with prune_ratio 0.2 + quant_bit 8, full-clip eval
      (decode fps through main_eval's own FPS loop, PSNR / MS-SSIM)
Widened rows, same geometry as --config, named in `config.workload`; informational lines:
  --branch_type ACB|RepVGG|DBB|ECB|NeRV_vanilla   the reference's other branch sets (folded online here; the stock
                                                  PyTorch leg runs them as the reference does, one conv per branch)
  --act gelu|relu|...                             the other activations (pre-activation mode + onr_act_map)
  --finetune_prune R                              the prune-then-finetune step (main_eval.py:446-507): global
                                                  magnitude masks at ratio R, masked-gradient FrameFitter

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

_COMMON = dict(embed='1.25_40', stem_dim_num='512_1', expansion=1, reduction=2, lower_width=96, branch_type='ERB',
               act='swish', loss_type='Fusion6', lr=5e-4, epochs=300, warmup_ratio=0.2, beta=0.5)
WORKLOADS = {
    "c1": dict(_COMMON, name="ERB Bunny-shaped synthetic 132x720x1280 (BASELINE configs[1])", kind="train",
               n_frames=132, H=720, W=1280, fc_hw_dim='9_16_26', strides=[5, 2, 2, 2, 2]),
    "c2": dict(_COMMON, name="ERB Bunny-shaped synthetic 132x720x1280, fc_hw_dim 9_16_112 NeRV-L width "
               "(BASELINE configs[2])", kind="train", n_frames=132, H=720, W=1280, fc_hw_dim='9_16_112',
               strides=[5, 2, 2, 2, 2]),
    "c3": dict(_COMMON, name="ERB UVG-shaped synthetic 600x1080x1920, strides 5 3 2 2 2, fc_hw_dim 9_16_26 "
               "(BASELINE configs[3])", kind="train", n_frames=600, H=1080, W=1920, fc_hw_dim='9_16_26',
               strides=[5, 3, 2, 2, 2]),
    "c4": dict(_COMMON, name="reparameterised single-branch decode of the configs[1] model, prune_ratio 0.2 + "
               "quant_bit 8, full-clip eval 132x720x1280 (BASELINE configs[4])", kind="decode", n_frames=132, H=720,
               W=1280, fc_hw_dim='9_16_26', strides=[5, 2, 2, 2, 2], prune_ratio=0.2, quant_bit=8, fit_epochs=3),
}
WORKLOAD = WORKLOADS["c1"]          # replaced by --config in main()


def geometry(w):
    """Per-block implicit-GEMM shapes (M = H*W pixels, N = Cnew*s^2, K = 9*Cin) and algorithmic MACs per frame
    (SURVEY.md 8d): stem + block convolutions + head; a training step is 3x the forward (fprop + dgrad + wgrad)."""
    fh, fw, fd = [int(x) for x in w['fc_hw_dim'].split('_')]
    stem_dim = int(w['stem_dim_num'].split('_')[0])
    macs = 80 * stem_dim + stem_dim * fh * fw * fd
    blocks, h, wd, c = [], fh, fw, fd
    fold_macs = 0
    for i, s in enumerate(w['strides']):
        cnew = int(c * w['expansion']) if i == 0 else max(c // w['reduction'], w['lower_width'])
        cout = cnew * s * s
        blocks.append(dict(M=h * wd, N=cout, K=9 * c, cin=c, cout=cout))
        macs += h * wd * cout * 9 * c
        fold_macs += 18 * cout * c * c + 9 * cout * cout * c
        h, wd, c = h * s, wd * s, cnew
    macs += h * wd * c * 3
    assert (h, wd) == (w['H'], w['W']), (h, wd)
    return dict(blocks=blocks, fwd_gflop=2e-9 * macs, step_gflop=6e-9 * macs, fold_fwd_gflop=2e-9 * fold_macs,
                last_gemm_gflop=2e-9 * blocks[-1]['M'] * blocks[-1]['N'] * blocks[-1]['K'])


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

    def __init__(self, gpu_index, period_s=0.005):
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
def host_threads():
    """All the host threads this process may use (torchrun exports OMP_NUM_THREADS=1, which would silently turn the
    CPU arm into a single-core run)."""
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def synthetic_frames_cpu(n, H, W):
    """Smooth synthetic frames built with plain torch on the host (the CPU / library arms do not import the package)."""
    import torch
    ys = torch.linspace(0, 1, H).view(1, 1, H, 1)
    xs = torch.linspace(0, 1, W).view(1, 1, 1, W)
    ph = torch.arange(n, dtype=torch.float32).view(n, 1, 1, 1) * 0.37
    ch = torch.arange(3, dtype=torch.float32).view(1, 3, 1, 1)
    img = 0.5 + 0.25 * torch.sin(6.0 * ys + 2.0 * ch + ph) * torch.cos(5.0 * xs - ch + 0.5 * ph)
    return (img.clamp(0, 1) * 255).round().div(255)


def oracle_step_loop(n_steps, warmup, device, sync=None):
    """`n_steps` iterations of reference main_train.py:229-254 written with the oracle's functions (forward, Fusion6,
    autograd backward, Adam, PSNR, MS-SSIM) on `device`.  Returns (seconds for n_steps, last loss)."""
    import torch
    from oracle import nerv_oracle as O
    w = WORKLOAD
    fh, fw, fd = [int(x) for x in w['fc_hw_dim'].split('_')]
    cfg = dict(fc_h=fh, fc_w=fw, fc_dim=fd, strides=w['strides'], sigmoid=False, act=w['act'])
    if w['branch_type'] in ('ACB', 'RepVGG', 'DBB', 'ECB'):
        cfg['explicit_branches'] = w['branch_type']     # as the reference runs them (model.py:541-565): no fold
    sd = {k: v.to(device) for k, v in O.random_state(cfg, w, seed=1, branch_type=w['branch_type']).items()}
    frames = synthetic_frames_cpu(2, w['H'], w['W']).to(device)
    state = {}
    total = 0.0
    loss = None
    for i in range(warmup + n_steps):
        pos = torch.tensor([(i % 2) / w['n_frames']])
        embed = O.pos_encoding(pos, 1.25, 40).to(device)            # PE on the CPU, then uploaded (main_train.py:234)
        if sync:
            sync()
        t0 = time.perf_counter()
        lr = O.lr_at(0, i, w['n_frames'], w['lr'], int(w['warmup_ratio'] * w['epochs']), w['epochs'])
        sd, state, loss, img, _ = O.train_step(sd, state, embed, frames[i % 2:i % 2 + 1], cfg, lr, i + 1)
        _ = O.psnr(img, frames[i % 2:i % 2 + 1])
        _ = O.ms_ssim(img, frames[i % 2:i % 2 + 1])
        if sync:
            sync()
        if i >= warmup:
            total += time.perf_counter() - t0
    return total, float(loss)


def oracle_decode_loop(n_frames, warmup, device, sync=None):
    """Decode (forward only) of a single-branch model with the oracle's functions: reference main_eval.py:753-762."""
    import torch
    from oracle import nerv_oracle as O
    w = WORKLOAD
    fh, fw, fd = [int(x) for x in w['fc_hw_dim'].split('_')]
    cfg = dict(fc_h=fh, fc_w=fw, fc_dim=fd, strides=w['strides'], sigmoid=False)
    sd = {k: v.to(device) for k, v in O.random_state(cfg, w, seed=1, deploy=True).items()}
    total = 0.0
    with torch.no_grad():
        for i in range(warmup + n_frames):
            embed = O.pos_encoding(torch.tensor([(i % w['n_frames']) / w['n_frames']]), 1.25, 40).to(device)
            if sync:
                sync()
            t0 = time.perf_counter()
            O.generator_forward(sd, embed, cfg)
            if sync:
                sync()
            if i >= warmup:
                total += time.perf_counter() - t0
    return total


def cpu_steps(n_steps, warmup, threads=None):
    """Times `n_steps` steps of the configured workload with the oracle port of the reference on host cores."""
    import torch
    torch.set_num_threads(threads or host_threads())
    if WORKLOAD['kind'] == 'decode':
        return oracle_decode_loop(n_steps, warmup, torch.device('cpu')), torch.get_num_threads()
    total, _ = oracle_step_loop(n_steps, warmup, torch.device('cpu'))
    return total, torch.get_num_threads()


def metric_of(w):
    if w['kind'] == 'decode':
        return "reparam decode frames/s (720p, prune 0.2 + quant 8)"
    extra = "" if w['act'] == 'swish' else ", act " + w['act']
    if w.get('finetune_prune'):
        extra += ", prune-then-finetune step at ratio {}".format(w['finetune_prune'])
    return "train frames/s ({}p {}{})".format(w['H'], w['branch_type'], extra)


def run_reference(opts):
    """CPU arm: the oracle port of the reference step on all host threads (rank 0 only under torchrun)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    w = WORKLOAD
    # bounded sample: a CPU step of this workload takes 0.2 - 2.5 s, so at most 30 timed steps (and 2 warm-up steps)
    # whatever --steps / --warmup ask for — the whole arm then ends within about a minute
    asked = (opts.steps, opts.warmup)
    opts.steps, opts.warmup = min(opts.steps, 30), min(opts.warmup, 2)
    total, cores = cpu_steps(opts.steps, opts.warmup)
    fps = opts.steps / total
    what = ("forward-only decodes of one frame (single-branch model)" if w['kind'] == 'decode' else
            f"{w['branch_type']} training steps of one frame (fwd, Fusion6, bwd, Adam, PSNR, MS-SSIM)")
    line = {
        "impl": "reference", "metric": metric_of(w), "value": fps, "unit": "frames/s",
        "n_gpus": opts.gpus, "steps": opts.steps, "warmup": opts.warmup, "ms_per_step": 1000.0 * total / opts.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": w['name'], "batch_per_gpu": 1, "note": "CPU arm: oracle port of the "
                   "reference step (reference is Python/PyTorch and /root/reference does not travel to the GPU box); "
                   f"bounded sample: {opts.steps} timed steps after {opts.warmup} warm-up (asked for {asked[0]} / {asked[1]})"},
        "cpu_baseline": {"value": fps, "unit": "frames/s", "cores": cores, "kind": "port",
                         "sample": f"{opts.steps} {what} at {w['H']}x{w['W']}"},
        "e2e": {"value": fps, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------------------------- GPU library arm
def torch_gpu_baseline(steps, warmup, kernels=True):
    """Stock PyTorch on the same B200, as a user of the reference would run it: the oracle's
    restatement of model.py / utils.py / main_train.py:229-254 on CUDA tensors with PyTorch's default TF32 convolutions
    and `cudnn.benchmark = True` (main_train.py:161).  Informational baseline: cuDNN / cuBLAS / ATen kernels only."""
    import torch
    import torch.nn.functional as F
    dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", "0")))
    torch.backends.cudnn.benchmark = True
    w = WORKLOAD
    sync = torch.cuda.synchronize
    if w['kind'] == 'decode':
        total = oracle_decode_loop(steps, warmup, dev, sync)
    else:
        total, _ = oracle_step_loop(steps, warmup, dev, sync)
    out = {"value": steps / total, "unit": "frames/s", "ms_per_step": 1000.0 * total / steps, "steps": steps,
           "what": "oracle restatement of the reference step on CUDA, TF32 convs (PyTorch default), "
                   "cudnn.benchmark=True, host-timed with synchronize per step",
           "torch": torch.__version__, "cudnn": torch.backends.cudnn.version()}
    if kernels and w['kind'] == 'train':
        # the block-4 convolution alone, cuDNN's best algorithm: TF32 (what the reference runs) and bf16 channels-last
        # (the same operand precision as conv_igemm_kernel / wgrad_igemm_kernel)
        g = geometry(w)['blocks'][-1]
        Hh, Ww = w['H'] // w['strides'][-1], w['W'] // w['strides'][-1]
        res = {}
        for name, dt, fmt in (("tf32_nchw", torch.float32, torch.contiguous_format),
                              ("bf16_nhwc", torch.bfloat16, torch.channels_last)):
            x = torch.randn(1, g['cin'], Hh, Ww, device=dev, dtype=dt).contiguous(memory_format=fmt)
            wt = torch.randn(g['cout'], g['cin'], 3, 3, device=dev, dtype=dt).contiguous(memory_format=fmt)
            gy = torch.randn(1, g['cout'], Hh, Ww, device=dev, dtype=dt).contiguous(memory_format=fmt)
            fns = {"fprop": lambda: F.conv2d(x, wt, None, 1, 1),
                   "dgrad": lambda: torch.ops.aten.convolution_backward(gy, x, wt, None, [1, 1], [1, 1], [1, 1], False,
                                                                        [0, 0], 1, [True, False, False]),
                   "wgrad": lambda: torch.ops.aten.convolution_backward(gy, x, wt, None, [1, 1], [1, 1], [1, 1], False,
                                                                        [0, 0], 1, [False, True, False])}
            for op, fn in fns.items():
                for _ in range(3):
                    fn()
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record()
                for _ in range(10):
                    fn()
                b.record()
                torch.cuda.synchronize()
                ms = a.elapsed_time(b) / 10
                res[f"{op}_{name}_ms"] = round(ms, 4)
                res[f"{op}_{name}_tflops"] = round(geometry(w)['last_gemm_gflop'] / ms, 1)
            del x, wt, gy
        out["block4_cudnn"] = res
    return out


def run_torch_gpu(opts):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    w = WORKLOAD
    r = torch_gpu_baseline(opts.steps, opts.warmup)
    line = {"impl": "torch-gpu", "metric": metric_of(w), "value": r["value"], "unit": "frames/s", "n_gpus": 1,
            "steps": opts.steps, "warmup": opts.warmup, "ms_per_step": r["ms_per_step"], "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "tf32", "data": "synthetic",
            "config": {"workload": w['name'], "batch_per_gpu": 1, "note": r["what"]},
            "gpu_library_baseline": r, "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------------------------- GPU arm
def _mark(msg):
    if os.environ.get("ONR_BENCH_VERBOSE"):
        print(f"[bench rank {os.environ.get('RANK', '0')}] {msg}", file=sys.stderr, flush=True)


def reference_api_loop(gen, pe, args, pinned_frames_u8, pinned_t, n_frames, steps):
    """The reference training iteration, written exactly as main_train.py:229-254, against the drop-in modules
    (`orepnerv.model.Generator`, `orepnerv.utils.*`, `orepnerv.optim.FusedAdam`): host float frames in (what the
    reference DataLoader hands over, 11 MB at 720p), `.cuda(non_blocking=True)`, forward, loss_fn, adjust_lr,
    zero_grad, backward, step, psnr_fn, msssim_fn, and one host read of the PSNR per step."""
    import copy
    import torch
    import torch.nn.functional as F
    from orepnerv.optim import FusedAdam
    from orepnerv.utils import adjust_lr, loss_fn, msssim_fn, psnr_fn
    model = copy.deepcopy(gen)
    model.train()
    PE = pe
    optimizer = FusedAdam(model.parameters(), betas=(args.beta, 0.999))
    a = argparse.Namespace(**vars(args))
    a.lw = 1.0
    data_size = n_frames
    loader = [(f.float().div(255).pin_memory(), t) for f, t in zip(pinned_frames_u8, pinned_t)]
    h2d = loader[0][0].numel() * 4 + 80 * 4

    def iteration(i, data, norm_idx):
        embed_input = PE(norm_idx)
        data, embed_input = data.cuda(non_blocking=True), embed_input.cuda(non_blocking=True)
        output_list = model(embed_input)
        target_list = [F.adaptive_avg_pool2d(data, x.shape[-2:]) for x in output_list]
        loss_list = [loss_fn(output, target, a) for output, target in zip(output_list, target_list)]
        loss_list = [loss_list[k] * (a.lw if k < len(loss_list) - 1 else 1) for k in range(len(loss_list))]
        loss_sum = sum(loss_list)
        adjust_lr(optimizer, 0, i, data_size, a)
        optimizer.zero_grad()
        loss_sum.backward()
        optimizer.step()
        psnr = psnr_fn(output_list, target_list)
        msssim_fn(output_list, target_list)
        return psnr

    for i in range(3):
        iteration(i, *loader[i % len(loader)]).tolist()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    last = None
    for i in range(steps):
        last = iteration(3 + i, *loader[i % len(loader)]).tolist()          # D2H read of the step's PSNR
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    return {"value": 1000.0 / ms, "unit": "frames/s", "ms_per_step": ms, "steps": steps, "h2d_bytes_per_step": h2d,
            "d2h_bytes_per_step": 4, "last_psnr": last[0][0],
            "what": "main_train.py:229-254 verbatim against orepnerv.model / utils / optim (autograd path, eager "
                    "launches, fp32 host frame uploaded every step)"}


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
                    expansion=w['expansion'], num_blocks=1, norm='none', act=w['act'], bias=True,
                    reduction=w['reduction'], conv_type='conv', stride_list=w['strides'], sin_res=True,
                    lower_width=w['lower_width'], sigmoid=False, deploy=False, branch_type=w['branch_type']).to(dev)
    n_frames = opts.frames or w['n_frames']
    clip = synthetic_clip(n_frames, w['H'], w['W'], device=dev)                 # uint8, resident in HBM
    spe = sharding.steps_per_epoch(n_frames, world)
    ft_kw, ft_info = {}, None
    if w.get('finetune_prune'):
        # the prune-then-finetune step (main_eval.py:239-507): global magnitude masks over the train-state tensors, the
        # ERB branch kernels frozen as in the reference, epoch numbering continued behind the 300 training epochs
        from orepnerv.main_eval import global_masks, train_state_prunable
        targets = train_state_prunable(gen)
        masks = global_masks([m.weight.detach() for _, m in targets], float(w['finetune_prune']))
        gm = {}
        with torch.no_grad():
            for (n, m), mask in zip(targets, masks):
                m.weight.mul_(mask)
                gm[n + '.weight'] = torch.zeros_like(mask) if ('.rbr_' in n) else mask
        ft_kw = dict(grad_masks=gm, epoch_offset=w['epochs'], epoch_mod=w['epochs'] + 100)
        ft_info = {"prune_ratio": w['finetune_prune'], "masked_tensors": len(gm),
                   "mask_zeros": int(sum(int((m == 0).sum()) for m in masks)),
                   "mask_total": int(sum(m.numel() for m in masks))}
    fit = FrameFitter(gen, pe, args, world_size=world, data_size=n_frames, steps_per_epoch=spe,
                      use_graph=not opts.no_graph, **ft_kw)
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
    geo = geometry(w)
    STEP_GFLOP, L4_GEMM_GFLOP = geo['step_gflop'], geo['last_gemm_gflop']
    last = len(ex.fprop) - 1
    for name, fn in (("conv_igemm_kernel<fprop,block4>", lambda: lib.onr_conv_plan_run(ex.fprop[last].handle, st)),
                     ("conv_igemm_kernel<dgrad,block4>", lambda: lib.onr_conv_plan_run(ex.dgrad[last].handle, st)),
                     ("wgrad_igemm_kernel<block4>", lambda: lib.onr_wgrad_plan_run(ex.wgrad[last].handle, st))):
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
    tpath = os.path.join(ROOT, "profiles", "r02_dram_traffic.json")
    if os.path.exists(tpath):
        with open(tpath) as f:
            traffic = json.load(f).get(dom)
    roofline = {"bound": "tensor", "kernel": dom, "achieved": achieved, "peak": peaks['bf16'], "unit": "TFLOP/s",
                "frac": achieved / peaks['bf16'], "traffic": traffic, "traffic_unit": "bytes/launch (ncu dram read+write)",
                "peak_source": peaks['source'] + " bf16 burst",
                "algorithmic_gflop_per_launch": L4_GEMM_GFLOP,
                "algorithmic_gflop_per_step": STEP_GFLOP, "fold_fwd_gflop_per_step": geo['fold_fwd_gflop'],
                "launch_ms": {k: round(v, 4) for k, v in kern.items()},
                "step_tflops": STEP_GFLOP / (ms / opts.steps), "step_frac_of_sustained": STEP_GFLOP / (ms / opts.steps) / peaks['bf16_sustained']}

    # ---------------- ERB fold, serialised on one stream: forward fold + pack of every block, fold backward of every block
    fold = {}
    for name, fn in (("fold_fwd_all_blocks", lambda: [ex.fold_pack_block(l) for l in range(ex.L)]),
                     ("fold_bwd_all_blocks", lambda: [ex.scatter_block_grads(l, fit.grads) for l in range(ex.L)])):
        for _ in range(2):
            fn()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(5):
            fn()
        b.record()
        torch.cuda.synchronize()
        fold[name + "_ms"] = round(a.elapsed_time(b) / 5.0, 4)
    fit.flat_grad.zero_()
    roofline["fold_serialised_ms"] = fold

    # ---------------- the reference's own loop (main_train.py:229-254) against the drop-in modules ----------------
    ref_api = reference_api_loop(gen, pe, args, pinned_frames, pinned_t, n_frames, max(10, min(opts.steps, 30))) \
        if world == 1 else None
    _mark("reference-API loop done")

    # ---------------- reparameterised decode (BASELINE metric's second figure; reference main_eval.py decode loop) ----
    decode = None
    if rank == 0:
        import copy
        dep = copy.deepcopy(gen)
        for blk in dep.layers:
            blk.switch_to_deploy()                     # ERB branches folded into one 3x3 conv per block
        dep.eval()
        with torch.no_grad():
            n_emb = min(8, n_frames)
            embeds = [pe(t_all[i:i + 1]) for i in range(n_emb)]
            for k in range(5):
                dep(embeds[k % n_emb])
            torch.cuda.synchronize()
            d0, d1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            n_dec = max(opts.steps, 20)
            d0.record()
            for k in range(n_dec):
                dep(embeds[k % n_emb])
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
            "metric": metric_of(w), "value": value, "unit": "frames/s", "n_gpus": world,
            "steps": opts.steps, "warmup": opts.warmup, "ms_per_step": ms / opts.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": w['name'], "batch_per_gpu": 1, "global_batch": world,
                       "parallelism": f"frame-sharded dp{world}", "l2": "per-step working set (~0.9 GB of "
                       "activations) exceeds the 126 MB L2; no explicit flush", "cuda_graph": fit.graph is not None,
                       "gradient_exchange": (fit.exchange + " (NCCL all-reduce captured in the step graph)")
                       if world > 1 else None,
                       "metrics_every_step": "PSNR + MS-SSIM (reference main_train.py:253-254)"},
            "clocks": clocks,
            "e2e": {"value": e2e, "unit": "frames/s", "ms_per_step": ms_e2e / opts.steps,
                    "h2d_bytes_per_step": int(pinned_frames[0].numel() + 4), "d2h_bytes_per_step": 32},
            "gpu_launches": int(per_step * opts.steps), "kernels_per_step": int(per_step),
            "roofline": roofline,
            "last_step": {"loss": out_last[0].item(), "psnr": out_last[4].item(), "msssim": out_last[5].item()},
            "finetune": ft_info,
            "decode": decode,
            "e2e_reference_api": ref_api,
        }
        if world == 1 and not opts.no_gpu_baseline:
            try:
                line["gpu_library_baseline"] = torch_gpu_baseline(8, 3)
            except Exception as exc:                                   # noqa: BLE001  (informational leg only)
                line["gpu_library_baseline"] = {"error": repr(exc)[:200]}
        if cpu_total is not None:
            line["cpu_baseline"] = {"value": 2 / cpu_total, "unit": "frames/s", "cores": cores, "kind": "port",
                                    "sample": f"2 {w['branch_type']} training steps of one {w['H']}x{w['W']} frame after 1 warm-up "
                                              "(oracle port)"}
        print(json.dumps(line), flush=True)
    if world > 1:
        # the step graph holds captured NCCL kernels: release it before the communicator goes away
        # (destroy_process_group otherwise blocks on the communicator the live graph still references)
        dist.barrier()
        fit.release_graph()
        del fit
        import gc
        gc.collect()
        torch.cuda.synchronize()
        dist.destroy_process_group()


def run_decode(opts):
    """BASELINE configs[4]: the reference eval flow (main_eval.py:551-827) on the reparameterised single-branch model —
    fit the ERB model briefly so that PSNR means something, `switch_to_deploy`, global prune 0.2, 8-bit quantisation,
    then decode the whole clip.  A "step" = the decode of one frame.  `value`: device-timed decode with the embeddings
    resident in HBM; `e2e`: main_eval's own FPS loop (host clock around synchronize, embedding computed per frame from a
    host index, PSNR read back) — the number the reference prints."""
    import copy
    import torch
    from orepnerv import _lib, main_eval
    from orepnerv.cli_common import FrameCache
    from orepnerv.model import Generator
    from orepnerv.trainer import FrameFitter
    from orepnerv.utils import PositionalEncoding

    w = WORKLOAD
    dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", "0")))
    torch.cuda.set_device(dev)
    lib = _lib.lib()
    args = make_args(1)
    torch.manual_seed(1)
    pe = PositionalEncoding(w['embed'])
    gen = Generator(embed_length=pe.embed_length, stem_dim_num=w['stem_dim_num'], fc_hw_dim=w['fc_hw_dim'],
                    expansion=w['expansion'], num_blocks=1, norm='none', act='swish', bias=True,
                    reduction=w['reduction'], conv_type='conv', stride_list=w['strides'], sin_res=True,
                    lower_width=w['lower_width'], sigmoid=False, deploy=False, branch_type=w['branch_type']).to(dev)
    n_frames = opts.frames or w['n_frames']
    cache = FrameCache(f"synthetic:{n_frames}x{w['H']}x{w['W']}", dev)
    # a short fit (fit_epochs passes over the clip, README schedule compressed) so the decode has something to show
    fa = argparse.Namespace(**vars(args))
    fa.epochs, fa.warmup = w['fit_epochs'], 0
    fit = FrameFitter(gen, pe, fa, data_size=n_frames, steps_per_epoch=n_frames, use_graph=True, with_msssim=False)
    for ep in range(w['fit_epochs']):
        for i in torch.randperm(n_frames, generator=torch.Generator().manual_seed(ep)).tolist():
            fit.step(cache.frames[i:i + 1], cache.t[i:i + 1])
    torch.cuda.synchronize()
    del fit
    dep = copy.deepcopy(gen)
    for blk in dep.layers:
        blk.switch_to_deploy()
    del gen
    ea = argparse.Namespace(prune_ratio=w['prune_ratio'], quant_bit=w['quant_bit'], quant_axis=0, print_freq=10 ** 9,
                            dump_images=False, outf=".")
    info = main_eval.prune_and_quantise(dep, ea, n_frames, (w['H'], w['W']))
    dep.eval()

    # ---- `value`: device-timed decode, embeddings resident, K frames after W warm-up
    with torch.no_grad():
        embeds = [pe(cache.t[i:i + 1]) for i in range(n_frames)]
        sampler = ClockSampler(dev.index or 0)
        sampler.start()                  # before the warm-up: NVML initialisation takes longer than a short timed region
        for k in range(max(opts.warmup, 50)):
            dep(embeds[k % n_frames])
        torch.cuda.synchronize()
        launches0 = lib.onr_launch_count()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for k in range(opts.steps):
            dep(embeds[k % n_frames])
        e1.record()
        torch.cuda.synchronize()
        clocks = sampler.stop()
        launches = lib.onr_launch_count() - launches0
    ms = e0.elapsed_time(e1) / opts.steps

    # ---- `e2e`: the reference's own full-clip loop (PSNR / MS-SSIM of every frame, 10 timed forwards per frame)
    res = main_eval.decode_clip(dep, pe, cache, ea, log_path=None, fwd_num=10, quiet=True)

    geo = geometry(w)
    peaks = measured_peaks()
    ex = dep.executor(1, False)
    st = torch.cuda.current_stream().cuda_stream
    last = len(ex.fprop) - 1
    for _ in range(3):
        lib.onr_conv_plan_run(ex.fprop[last].handle, st)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(10):
        lib.onr_conv_plan_run(ex.fprop[last].handle, st)
    b.record()
    torch.cuda.synchronize()
    kms = a.elapsed_time(b) / 10.0
    achieved = geo['last_gemm_gflop'] / kms
    roofline = {"bound": "tensor", "kernel": "conv_igemm_kernel<fprop_infer,block4>", "achieved": achieved,
                "peak": peaks['bf16'], "unit": "TFLOP/s", "frac": achieved / peaks['bf16'], "traffic": None,
                "peak_source": peaks['source'] + " bf16 burst", "algorithmic_gflop_per_launch": geo['last_gemm_gflop'],
                "launch_ms": {"conv_igemm_kernel<fprop_infer,block4>": round(kms, 4)},
                "frame_tflops": geo['fwd_gflop'] / ms, "frame_frac_of_sustained": geo['fwd_gflop'] / ms / peaks['bf16_sustained']}
    cpu_total, cores = cpu_steps(3, 1) if not opts.no_cpu_baseline else (None, None)
    line = {
        "metric": metric_of(w), "value": 1000.0 / ms, "unit": "frames/s", "n_gpus": 1, "steps": opts.steps,
        "warmup": opts.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "bf16", "data": "synthetic",
        "config": {"workload": w['name'], "batch_per_gpu": 1, "prune_quant": info.strip().split("\n"),
                   "fit": f"{w['fit_epochs']} epochs of FrameFitter steps before deploy (random init otherwise)",
                   "l2": "per-frame working set (~0.35 GB of activations) exceeds the 126 MB L2; no explicit flush"},
        "clocks": clocks,
        "e2e": {"value": res['fps'], "unit": "frames/s", "fps_first_frame": res['fps_first_frame'],
                "frames": res['frames'], "h2d_bytes_per_step": 4, "d2h_bytes_per_step": 8,
                "what": "main_eval.decode_clip: the reference FPS loop (main_eval.py:738-827), 10 forwards per frame, "
                        "host clock around synchronize, PSNR + MS-SSIM of every frame"},
        "quality": {"psnr": res['psnr'], "msssim": res['msssim']},
        "gpu_launches": int(launches), "kernels_per_step": launches / opts.steps, "roofline": roofline,
    }
    if cpu_total is not None:
        line["cpu_baseline"] = {"value": 3 / cpu_total, "unit": "frames/s", "cores": cores, "kind": "port",
                                "sample": f"3 forward-only decodes of one {w['H']}x{w['W']} frame after 1 warm-up "
                                          "(oracle port)"}
    if not opts.no_gpu_baseline:
        try:
            line["gpu_library_baseline"] = torch_gpu_baseline(20, 5)
        except Exception as exc:                                       # noqa: BLE001
            line["gpu_library_baseline"] = {"error": repr(exc)[:200]}
    print(json.dumps(line), flush=True)


def _arm_watchdog(seconds):
    """A wedged GPU wait must not hang the harness: dump every Python stack and exit non-zero."""
    import faulthandler
    faulthandler.enable()
    faulthandler.dump_traceback_later(seconds, exit=True)


def main():
    _arm_watchdog(int(os.environ.get("ONR_BENCH_WATCHDOG_S", "900")))
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference", "torch-gpu"])
    ap.add_argument("--config", default="c1", choices=sorted(WORKLOADS),
                    help="BASELINE.json configuration: c1 S720 (default), c2 L720, c3 U1080, c4 prune+quant decode")
    ap.add_argument("--no-gpu-baseline", action="store_true", help="skip the stock-PyTorch-on-GPU leg")
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--frames", type=int, default=0, help="clip length override (profiling runs only)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--branch_type", default=None, choices=["NeRV_vanilla", "ERB", "ACB", "RepVGG", "DBB", "ECB"])
    ap.add_argument("--act", default=None)
    ap.add_argument("--finetune_prune", type=float, default=0.0)
    opts = ap.parse_args()
    global WORKLOAD
    WORKLOAD = dict(WORKLOADS[opts.config])
    if opts.branch_type and opts.branch_type != WORKLOAD['branch_type']:
        WORKLOAD['branch_type'] = opts.branch_type
        WORKLOAD['name'] = WORKLOAD['name'].replace("ERB ", opts.branch_type + " ", 1) + \
            f" [branch_type {opts.branch_type}: SURVEY.md 8f-4, not a BASELINE config]"
    if opts.act and opts.act != WORKLOAD['act']:
        WORKLOAD['act'] = opts.act
        WORKLOAD['name'] += f" [act {opts.act}: SURVEY.md 8f-4, not a BASELINE config]"
    if opts.finetune_prune:
        WORKLOAD['finetune_prune'] = opts.finetune_prune
        WORKLOAD['name'] += f" [prune-then-finetune step, prune_ratio {opts.finetune_prune}: SURVEY.md 8f-3]"
    if opts.impl == "reference":
        run_reference(opts)
    elif opts.impl == "torch-gpu":
        run_torch_gpu(opts)
    else:
        opts.warmup = max(opts.warmup, 3)
        if WORKLOAD['kind'] == 'decode':
            run_decode(opts)
        else:
            run_ours(opts)


if __name__ == "__main__":
    main()
