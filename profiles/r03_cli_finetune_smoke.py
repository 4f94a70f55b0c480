"""Smoke of the prune-then-finetune command line (reference README workflow, main_eval.py --finetune) at toy size:
main_train writes model_latest.pth, main_eval --prune_ratio 0.3 --finetune --finetune_epochs 2 --quant_bit 8 prunes the
train-state model, fine-tunes, deploys, quantises and decodes.  Run from the repo root on a B200."""
import os
import sys
import tempfile

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, ROOT)
from orepnerv import main_eval, main_train  # noqa: E402

os.chdir(tempfile.mkdtemp())
for bt in ("ERB", "NeRV_vanilla"):
    flags = ("-e 3 --lr 0.002 -b 1 --embed 1.25_40 --stem_dim_num 64_1 --fc_hw_dim 3_4_12 --expansion 1 --reduction 2 "
             "--lower_width 8 --strides 3 2 --single_res --loss Fusion6 --warmup 0.2 --lr_type cosine --norm none "
             f"--act swish --branch_type {bt} --outf toy --suffix {bt} --dataset synthetic:6x18x24 --eval_freq 1 -p 100"
             ).split()
    main_train.main(flags + ['--overwrite'])
    psnr, ms = main_eval.main(flags + ['--eval_only', '--prune_ratio', '0.3', '--quant_bit', '8', '--finetune',
                                       '--finetune_epochs', '2'])
    out = os.path.join('result', 'toy', bt)
    files = sorted(os.listdir(out))
    assert 'finetune_e2_pr0.30_q8.txt' in files and 'only_prune0.30_quant8.txt' in files, files
    print(bt, 'finetune CLI ok: psnr', round(psnr, 2), files)
