"""Full-size GPU parity (B200): BASELINE configs[1..3] geometries at their real resolution against the fp32 GPU
oracle (oracle/nerv_oracle.py on CUDA tensors, TF32 off — SURVEY.md 8c-iv), plus the north_star convergence gate.

Stated tolerances (bf16 tensor-core operands, fp32 accumulate; the reference's own GPU convolutions are TF32):
  * decoded image            rel-L2 <= 5e-4  (measured 3.7e-5 .. 4.3e-5, profiles/r02_parity_fullsize.jsonl)
  * Fusion6 loss             |d|    <= 5e-5  (measured <= 2.1e-6)
  * every parameter gradient rel-L2 <= 1.5e-2 of that gradient's norm (measured <= 7.0e-3)
  * folded ERB kernels       rel-L2 <= 1e-5  (fp32 gate of north_star; fp64 oracle)
  * final eval PSNR after 30 epochs on the 132-frame clip of configs[1]: not more than 0.1 dB below the oracle and
    within 0.3 dB either way (measured -0.052 dB and +0.153 dB in two runs; see the test's docstring)
"""
import pytest
import torch

import fullsize_util as U      # tests/ is on sys.path (pytest rootdir-less "prepend" import mode)

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def dev():
    assert torch.cuda.is_available(), "GPU tests need a B200"
    from orepnerv import _lib
    _lib.lib()
    return torch.device("cuda:0")


@pytest.mark.parametrize("name", ["S720", "L720", "U1080"])
def test_one_step_parity_full_size(dev, name):
    res = U.one_step_parity(name, dev)
    U.record("one_step_parity", res)
    print({k: v for k, v in res.items() if k != "grad_rel_l2"})
    assert res["img_rel_l2"] <= 5e-4, res["img_rel_l2"]
    assert abs(res["loss"] - res["loss_ref"]) <= 5e-5
    assert abs(res["psnr"] - res["psnr_ref"]) <= 0.005
    assert res["grad_rel_l2_max"] <= 1.5e-2, (res["grad_rel_l2_worst"], res["grad_rel_l2_max"])
    assert res["fold_rel_l2_max"] <= 1e-5, res["fold_rel_l2_max"]


def test_one_step_parity_full_size_vanilla(dev):
    res = U.one_step_parity("S720", dev, branch_type="NeRV_vanilla")
    U.record("one_step_parity", res)
    assert res["img_rel_l2"] <= 5e-4 and abs(res["loss"] - res["loss_ref"]) <= 5e-5
    assert res["grad_rel_l2_max"] <= 1.5e-2, (res["grad_rel_l2_worst"], res["grad_rel_l2_max"])


def test_convergence_gate_s720(dev):
    """north_star gate 3: same frames, seed and frame order, fixed epoch count -> final PSNR against the fp32 oracle.
    BASELINE configs[1] itself: the 132-frame 720p clip, README recipe (main_train.py:222-267), 30 epochs (ours 5 s,
    oracle 130 s on a B200).  Measured deltas (ours - oracle) over independent runs: -0.052 dB and +0.153 dB
    (profiles/r02_convergence_S720_132f_30e*.json); the fp32 oracle moved by 0.04 dB between its own two runs and
    neither side is deterministic (cuDNN / atomics), so the gate is: NOT WORSE than the oracle by more than 0.1 dB, and
    within 0.3 dB either way, with the two training curves tracking each other all along.
    A reduced clip is NOT a usable gate: with 12 frames the fit goes through a phase transition whose onset is chaotic
    — three identical runs of the fp32 oracle itself ended at 34.9, 22.6 and 34.7 dB
    (profiles/r02_convergence_S720_12f_60e_3repeats.json); averaging over the 132 frames removes most of that."""
    res = U.convergence_run(dev, "S720", n_frames=132, epochs=30)
    U.record("convergence", res)
    print({k: v for k, v in res.items() if not isinstance(v, list)})
    assert res["ours_train_psnr"][-1] > res["ours_train_psnr"][0] + 10.0         # it does fit the clip
    assert res["delta_eval_psnr"] >= -0.1, res["delta_eval_psnr"]
    assert abs(res["delta_eval_psnr"]) <= 0.3, res["delta_eval_psnr"]
    assert abs(res["delta_train_psnr_last"]) <= 0.3, res["delta_train_psnr_last"]
    worst = max(abs(a - b) for a, b in zip(res["ours_train_psnr"][5:], res["oracle_train_psnr"][5:]))
    assert worst <= 0.6, worst
