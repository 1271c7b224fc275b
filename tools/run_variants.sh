#!/bin/bash
# dev tool: time every prebuilt library variant under variants/ on the north-star shape (two rounds, same box)
cp anml_b200/lib/libanml_b200.so /tmp/lib_orig.so
for round in 1 2; do
for f in variants/*.so; do
  cp $f anml_b200/lib/libanml_b200.so
  echo -n "$f " ; timeout 100 python tools/time_fused.py 1e8 64 1 2>&1 | tail -1 | cut -c1-90
done
done
cp /tmp/lib_orig.so anml_b200/lib/libanml_b200.so
