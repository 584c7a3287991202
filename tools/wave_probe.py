"""Wave quantisation probe: the same conv at patch counts around a whole number of waves on the 74 CTA pairs
(32x32, 256 -> 256: 4 pair tiles per patch; 16x16, 512 -> 512: 2; 8x8, 768 -> 768 with 192-wide tiles: 1)."""
import os
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(REPO, "tools"))
import tc_probe  # noqa: E402  (runs its default shape list on import: harmless, short)

for (C, H) in ((256, 32), (512, 16), (768, 8)):
    for P in (37, 55, 56, 64, 74, 92, 111):
        tc_probe.run(P, C, C, H, 9, 1, iters=30)
