"""Generate tests/golden/eval_ensemble.npz by executing the UNMODIFIED reference model (authoring container only).

Run:  PYTHONDONTWRITEBYTECODE=1 python oracle/gen_golden_eval.py [n_checkpoints]
Config 5 of BASELINE.json: 10 checkpoints of the reference aadensenet121 (`DenseNet(32,
(6,12,24,16), 64, num_classes=5, attn_params=...)`, chexpert.py:475-476) initialised under seeds 0..9, the fixed
synthetic 234-image evaluation set (chexpert_b200.evaluate.synthetic_radiographs_u8, seed 3), targets seed 4.  Stored:
every checkpoint's raw logits from the reference in eval mode (chexpert.py:196-214), the ensemble mean
(chexpert.py:233), sklearn's per-class AUROC of the mean and of every checkpoint (chexpert.py:130-135), a checksum of
every checkpoint's parameters (so the GPU test can prove it rebuilt the same weights from the seed), a checksum of the
images, and the transition-3 attention map the reference stashes for image 0 of checkpoint 0 (attn_aug_conv.py:87).
"""
import contextlib
import io
import os
import sys
import time

import numpy as np
import torch

sys.dont_write_bytecode = True
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, '..'))
sys.path.insert(0, '/root/reference/models')
import attn_aug_conv as ref  # noqa: E402  (the reference itself)
from chexpert_b200.evaluate import synthetic_radiographs_u8, synthetic_eval_targets, normalise_u8  # noqa: E402  (input generators only)
from oracle.aaconv_oracle import auroc_per_class, ensemble_mean  # noqa: E402

OUT = os.path.join(HERE, '..', 'tests', 'golden', 'eval_ensemble.npz')
N_IMAGES, BATCH = 234, 16      # explore_data.ipynb:39 validation-set size; chexpert.py:50 default batch


def param_checksum(sd):
    return float(sum(v.double().sum() for v in sd.values())), float(sum(v.double().abs().sum() for v in sd.values()))


def main():
    n_ckpt = int(sys.argv[1]) if len(sys.argv) > 1 else 10
    torch.set_num_threads(os.cpu_count())
    images = synthetic_radiographs_u8(N_IMAGES, 320, seed=3)
    targets = synthetic_eval_targets(N_IMAGES, seed=4)
    logits, sums, attn0 = [], [], None
    for seed in range(n_ckpt):
        t0 = time.time()
        torch.manual_seed(seed)
        with contextlib.redirect_stdout(io.StringIO()):
            net = ref.DenseNet(32, (6, 12, 24, 16), 64, num_classes=5,
                               attn_params={'k': 0.2, 'v': 0.1, 'nh': 8, 'relative': True, 'input_dims': (320, 320)})
        net.eval()
        sums.append(param_checksum(net.state_dict()))
        outs = []
        with torch.no_grad():
            for i in range(0, N_IMAGES, BATCH):
                outs.append(net(normalise_u8(images[i:i + BATCH])))
                if seed == 0 and i == 0:
                    attn0 = net.features.transition3.conv.weights[0].clone().numpy()
        logits.append(torch.cat(outs, 0))
        print(f'checkpoint {seed}: {time.time() - t0:.1f}s  logit std over images {logits[-1].std(0).tolist()}', flush=True)
    mean = ensemble_mean(logits)
    rec = {'per_model': torch.stack(logits, 0).numpy(), 'mean': mean.numpy(), 'targets': targets.numpy(),
           'auroc_mean': np.array(auroc_per_class(mean, targets)),
           'auroc_per_model': np.array([auroc_per_class(z, targets) for z in logits]),
           'param_checksums': np.array(sums), 'image_checksum': np.array([int(images.long().sum()), int(images[::7, ::5, ::3].long().sum())]),
           'attn_t3_img0_ckpt0': attn0}
    np.savez_compressed(OUT, **rec)
    print('wrote', OUT, {k: v.shape for k, v in rec.items()})
    print('AUROC of the mean', rec['auroc_mean'])


if __name__ == '__main__':
    main()
