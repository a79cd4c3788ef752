#!/bin/bash
# Same-box A/B of two builds of the library on the C3 training step.  usage: bash scripts/ab_lib_train.sh /path/to/other.so
for i in 1 2 3; do
  for lib in "" "$1"; do
    echo -n "lib ${lib:-default}: "; TEXPOSE_B200_LIB=$lib python scripts/step_timeline.py 20 2>&1 | grep "C3 train"
  done
done
