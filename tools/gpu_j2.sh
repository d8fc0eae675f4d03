#!/bin/bash
cd "$(dirname "$0")/.."
P="timeout 120 python tools/train_profile.py --only NOTHING --steps 2"
for v in "--overlap-prep 0 --opt gn_bwd_sx=0" "--overlap-prep 1 --opt gn_bwd_sx=0" "--overlap-prep 0 --opt gn_bwd_sx=1" "--overlap-prep 1 --opt gn_bwd_sx=1" "--overlap-prep 0 --opt gn_bwd_sx=0"; do
  echo "=== $v"; $P $v 2>&1 | grep "graph-replayed"
done
