from .discriminator import Discriminator  # noqa: F401
from .edgegan import EdgeGAN  # noqa: F401
from .encoder import Encoder  # noqa: F401
from .generator import Generator  # noqa: F401
