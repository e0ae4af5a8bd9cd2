#!/usr/bin/env python
"""Drop-in for the reference's trainer_LRID.py entry point (eval modes), running on pnnp_b200:
    python trainer_LRID.py -f runfiles/IMX686/PNNP.yml --mode evaltest"""
from pnnp_b200.trainer import main_lrid

if __name__ == '__main__':
    main_lrid()
