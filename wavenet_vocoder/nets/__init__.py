from pytorchwavenetvocoder_b200.nets import *  # noqa
