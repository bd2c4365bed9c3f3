"""ORACLE (test infrastructure) — re-export of the seeded synthetic checkpoint / input generators.

The generators themselves live in tools/synthetic.py (they are workload generators, not part of the
oracle: bench.py uses them for its synthetic `state_dict`s without importing anything from oracle/).
"""
from tools.synthetic import *  # noqa: F401,F403
from tools.synthetic import (IMAGENET_MEAN, IMAGENET_STD, _perturb_and_calibrate, swin_model,  # noqa: F401
                             synthetic_images, torchvision_model, torchvision_state_dict, vit_state_dict)
