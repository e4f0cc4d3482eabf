#!/bin/bash
# round 2, call S: end-to-end training curves through the reference's CLI defaults
mkdir -p gpurun_out/r2s
timeout 900 python tools/train_curves.py > gpurun_out/r2s/train_curves.log 2>&1; echo "curves rc=$?"
tail -5 gpurun_out/r2s/train_curves.log
