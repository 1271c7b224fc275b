cp anml_b200/lib/libanml_b200.so /tmp/lib_orig.so
for r in 1 2; do for f in variants/*.so; do cp $f anml_b200/lib/libanml_b200.so; echo -n "$f "; timeout 100 python tools/time_spline.py 5e7 2>&1 | tail -1 | cut -c60-175; done; done
cp /tmp/lib_orig.so anml_b200/lib/libanml_b200.so
