#!/usr/bin/env python
"""Drop-in for the reference's trainer_SID.py entry point (eval modes), running on pnnp_b200:
    python trainer_SID.py -f runfiles/SonyA7S2/PNNP.yml --mode evaltest"""
from pnnp_b200.trainer import main_sid

if __name__ == '__main__':
    main_sid()
