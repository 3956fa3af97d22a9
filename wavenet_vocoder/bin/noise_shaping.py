#!/usr/bin/env python
from pytorchwavenetvocoder_b200.bin.noise_shaping import *  # noqa
from pytorchwavenetvocoder_b200.bin.noise_shaping import main

if __name__ == "__main__":
    main()
