#!/bin/bash
mkdir -p gpurun_out
python tools/probe_farfield.py 512 2>&1 | tail -2
LSF_LIB_PATH=$PWD/variants/nopf.so python tools/probe_farfield.py 512 2>&1 | tail -2
LSF_LIB_PATH=$PWD/variants/pf1.so python tools/probe_farfield.py 512 2>&1 | tail -2
