#!/bin/bash
# Raw-chunk policy of the host-narrow upload (see upload_narrow_host): adaptive (-1) vs fixed patterns.
mkdir -p gpurun_out
out=gpurun_out/upload_policy.txt
: > $out
nproc >> $out
for re in -1 0 4 5 6; do
  echo "## ZB_UPLOAD_RAW_EVERY=$re" >> $out
  ZB_UPLOAD_RAW_EVERY=$re timeout 300 python tools/upload_bench.py 30 2>&1 | grep pinned >> $out
done
for g in 48 53 58; do
  echo "## adaptive ZB_UPLOAD_PCIE_GBS=$g" >> $out
  ZB_UPLOAD_PCIE_GBS=$g timeout 300 python tools/upload_bench.py 30 2>&1 | grep pinned >> $out
done
cat $out
