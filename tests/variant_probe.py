"""Helper of tests/test_gpu_variants.py (not a test module): a few fused train steps on cuda:0 under whatever ADER_B200_*
switches the environment carries (they are read once per process), then ONE line `DIGEST <sha256>` over the step losses, the
parameters, both Adam slots, the gradient of the last step and the step counter."""
import hashlib
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from ader_b200.model import Ader  # noqa: E402


def main():
    args = type("Args", (), dict(hidden_units=150, maxlen=50, num_blocks=2, num_heads=1, random_seed=0, lr=1e-3,
                                 dropout_rate=0.0, disable_distillation=False, loss_impl="tc"))()
    m = Ader(800, args, init_seed=0)
    assert m.step_impl == "dag" and m.encoder_impl == "tc", (m.step_impl, m.encoder_impl)
    m.update_loss(0.9)
    rng = np.random.RandomState(33)
    B, Me, V, Vp, E, L = 96, 32, 700, 650, 48, 50
    teacher = torch.randn(E, Vp, device=m.device, generator=torch.Generator(device=m.device).manual_seed(2))
    h = hashlib.sha256()
    for it in range(4):
        ids = np.zeros((B + Me, L), np.int32)
        for r in range(B + Me):
            n = int(rng.randint(1, 30))
            hot = r % 3 == 0            # a third of the rows draw from four items: segments that span several scatter windows
            ids[r, L - n:] = rng.randint(1, 5 if hot else Vp + 1, n)
        pos = rng.randint(1, V + 1, B).astype(np.int32)
        rows = rng.randint(0, E, Me).astype(np.int32)
        loss = m.train_step(ids, pos, V, 1e-3, 0.3, exemplar_logits=teacher, teacher_rows=rows, n_tokens=int((ids != 0).sum()))
        h.update(np.float32(loss.item()).tobytes())
    torch.cuda.synchronize()
    for t in (m.theta, m.adam_m, m.adam_v, m.grad, m.adam_state):
        h.update(t.detach().cpu().numpy().tobytes())
    print("DIGEST", h.hexdigest())


if __name__ == "__main__":
    main()
