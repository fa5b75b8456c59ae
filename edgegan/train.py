"""`python -m edgegan.train` (reference edgegan/train.py:138) -> edgegan_b200.train, same flags."""
from edgegan_b200.train import *  # noqa: F401,F403
from edgegan_b200.train import main

if __name__ == "__main__":
    main()
