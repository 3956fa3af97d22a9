from .wavenet import *  # noqa
