# A/B of kernel build variants: every gecco_b200/lib_<name>.so next to the production library is timed on config 2
# and the sparse shape (tools/path_time.py), production first.  B200 only.
for lib in libgecco_crf_b200.so $(cd gecco_b200 && ls lib_*.so 2>/dev/null); do
  echo "== $lib"
  GCRF_LIB_NAME=$lib timeout 300 python tools/path_time.py config2 sparse config4 2>&1 | grep -E "fused"
done
