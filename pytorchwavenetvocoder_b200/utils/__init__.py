from .utils import *  # noqa
from .utils import read_wav, write_wav  # noqa
