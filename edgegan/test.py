"""`python -m edgegan.test` (reference edgegan/test.py:130) -> edgegan_b200.test, same flags."""
from edgegan_b200.test import *  # noqa: F401,F403
from edgegan_b200.test import main

if __name__ == "__main__":
    main()
