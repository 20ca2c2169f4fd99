for v in r1 "" t256r t256r0 t512s0 t384r0 t384s0; do
  if [ -z "$v" ]; then unset NDTB_LIB; else export NDTB_LIB=$PWD/ndt_feature_graph_b200/lib/variants/libndtb_$v.so; fi
  python scripts/bench_match.py 592 3 2>&1 | tail -1 | cut -c1-330
done
