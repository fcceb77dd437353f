"""Teacher-student soft-DTW alignment loss: the consumer of kernel (2) that wav2vec2/lib.py:130,184-191 sketches
(``soft_dtw = SoftDTW(use_cuda=True, gamma=1.5)``; the clean copy's posteriors are the target sequence, every
augmented copy is aligned to them; BASELINE.json configs[3] sizes it at 4096 x 4096 frames, batch 8).

    pseudo_targets = logits[-1].detach().unsqueeze(0).repeat(B, 1, 1)      # teacher = clean branch
    predictions    = logits[:-1]                                            # students = augmented branches
    loss           = soft_dtw(pseudo_targets, predictions).mean()

The distance matrix is the reference module's squared Euclidean one (soft_dtw_cuda.py:319-329) formed with one
batched GEMM instead of the [B,N,M,d] expansion; forward/backward of the alignment run in dae_softdtw_fwd/bwd.
"""
import torch

from .soft_dtw_cuda import SoftDTW


def teacher_student_softdtw_loss(posteriors: torch.Tensor, num_negatives: int = None, gamma: float = 1.5,
                                 normalize: bool = False, bandwidth=None, module: SoftDTW = None):
    """posteriors [B+1, T', C]: rows 0..B-1 are the augmented (student) branches, the last row is the clean
    (teacher) branch, as in the adapt batch [aug..., clean] (lcasr/lib.py:539-541, wav2vec2/lib.py:184-191).
    Returns the mean soft-DTW value over the students; gradients flow to the student rows only."""
    B = posteriors.shape[0] - 1 if num_negatives is None else int(num_negatives)
    if B < 1 or posteriors.shape[0] < B + 1:
        raise ValueError("need at least one student row and the teacher row")
    sd = module if module is not None else SoftDTW(use_cuda=True, gamma=gamma, normalize=normalize, bandwidth=bandwidth)
    teacher = posteriors[-1].detach().unsqueeze(0).repeat(B, 1, 1)
    students = posteriors[:B]
    return sd(teacher, students).mean()
