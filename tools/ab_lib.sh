#!/bin/bash
# A/B of two builds of libcmu_b200.so on the same box: tools/ab/libcmu_b200.so (older build, copied there by hand) against the
# in-tree one, alternating.  usage: tools/ab_lib.sh <python script and args>
set -e
ROOT=$(cd "$(dirname "$0")/.." && pwd)
NEW=$ROOT/contrastive_masked_unet_b200/libcmu_b200.so
OLD=$ROOT/tools/ab/libcmu_b200.so
cp "$NEW" /tmp/libcmu_new.so
for round in 1 2; do
  echo "== old build (round $round)"; cp "$OLD" "$NEW"; python "$@"
  echo "== new build (round $round)"; cp /tmp/libcmu_new.so "$NEW"; python "$@"
done
