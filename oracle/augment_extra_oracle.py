"""CPU restatement of the reference's secondary augmentations (test infrastructure only).

frame_shuffle     lcasr/lib.py:81-84:   spec[:, :, perm_t] then spec[:, perm_f, :]
add_random_noise  lcasr/lib.py:379-382: spec + normal(0, spec.std(), size) * noise_factor

Both take the randomness as explicit descriptors (the permutations; the standard-normal field z, with
torch.normal(0, s, size) == randn(size) * s bitwise for the same generator state — asserted in
tests/test_oracle_pins.py).  PINNED by tests/golden/loop_toy.npz: outputs of the reference's own functions under a
fixed torch seed (tests/golden/make_loop_golden.py).
"""
import numpy as np


def frame_shuffle(spec, perm_t=None, perm_f=None):
    """spec [B,F,T] numpy."""
    if perm_t is not None:
        spec = spec[:, :, np.asarray(perm_t)]
    if perm_f is not None:
        spec = spec[:, np.asarray(perm_f), :]
    return spec


def add_random_noise(spec, z, noise_factor):
    """spec, z [B,F,T] float32 numpy; std = unbiased standard deviation in float64, rounded once to float32."""
    if noise_factor == 0:
        return spec
    x = spec.astype(np.float64)
    std = np.float32(np.sqrt(((x - x.mean()) ** 2).sum() / (x.size - 1)))
    noise = (z.astype(np.float32) * std).astype(np.float32)
    return (spec + (noise * np.float32(noise_factor)).astype(np.float32)).astype(np.float32)
