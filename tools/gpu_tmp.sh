#!/bin/bash
set -u
for i in 1 2; do
CERVIX_BN1_IN_DGRAD=1 python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-classifier 2>&1 | tail -1 | cut -c1-200
python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-classifier 2>&1 | tail -1 | cut -c1-200
done
