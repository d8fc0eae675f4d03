#!/bin/bash
cd "$(dirname "$0")/.."
P="timeout 120 python tools/train_profile.py --only NOTHING --steps 2"
for v in "--fuse-gn-bwd 0" "--fuse-gn-bwd 0 --opt gn_bwd_chunk=40" "--fuse-gn-bwd 0 --opt gn_bwd_chunk=64" "--fuse-gn-bwd 0 --opt gn_bwd_chunk=111" "--fuse-gn-bwd 1"; do
  echo "=== $v"; $P $v 2>&1 | grep "graph-replayed"
done
