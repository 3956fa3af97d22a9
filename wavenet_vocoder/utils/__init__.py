from pytorchwavenetvocoder_b200.utils import *  # noqa
from pytorchwavenetvocoder_b200.utils import read_wav, write_wav  # noqa
