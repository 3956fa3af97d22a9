#!/usr/bin/env python
from pytorchwavenetvocoder_b200.bin.train import *  # noqa
from pytorchwavenetvocoder_b200.bin.train import main

if __name__ == "__main__":
    main()
