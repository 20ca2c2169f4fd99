# A/B of match-kernel variants built by scripts/build_variants.sh: match time of one 592-pair batch, then the bench value
for v in "" ${VARIANTS:-t192b2 t128b3 t256b2}; do
  if [ -z "$v" ]; then unset NDTB_LIB; else export NDTB_LIB=$PWD/ndt_feature_graph_b200/lib/variants/libndtb_$v.so; fi
  echo "== variant '${v:-default}'"
  python scripts/bench_match.py 592 3 2>&1 | tail -1 | cut -c1-330
  python bench.py --no-cpu --no-extra --no-e2e --steps 6 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('value', d['value'], 'ms/step', d['ms_per_step'], 'match_ms', d['roofline']['launch_ms'], 'one lane', d['roofline']['step_ms_one_lane'])"
done
