#!/bin/bash
# per-stage device times of one query shape (development aid): events around one stage at a time
# usage: scripts/stage_times.sh NQ [D K N]
NQ=${1:-256}; D=${2:-384}; K=${3:-5}; N=${4:-1000000}
for st in 1 0 4 5; do
  echo -n "stage $st: "
  B2R_TIME_STAGE=$st python scripts/quick_gemm.py $NQ $D $K 20 $N 2>&1 | sed 's/.*| scoring kernel/kernel/; s/->.*|/|/'
done
