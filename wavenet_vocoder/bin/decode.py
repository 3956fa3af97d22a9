#!/usr/bin/env python
from pytorchwavenetvocoder_b200.bin.decode import *  # noqa
from pytorchwavenetvocoder_b200.bin.decode import main

if __name__ == "__main__":
    main()
