"""ader_b200 -- B200-native (sm_100a) implementation of ADER's per-period training,
evaluation and exemplar-selection path (reference: doublemul/ADER).

Host code is Python and mirrors the reference's own objects (`Ader` / `Ewc` model surface,
`DataLoader`, `Sampler`, `Evaluator`, `ExemplarGenerator`, the `main.py` CLI); all device work is
hand-written CUDA behind the C ABI of ``include/ader_b200.h`` (``libader_b200.so``), exposed as
``torch.ops.ader_b200.*``.  There is no CPU fallback and no other backend.
"""
from .params import Hyper, ParamLayout  # noqa: F401

__all__ = ["Hyper", "ParamLayout"]
