for v in "" p1024 p16384; do
  if [ -z "$v" ]; then unset NDTB_LIB; else export NDTB_LIB=$PWD/ndt_feature_graph_b200/lib/variants/libndtb_$v.so; fi
  timeout 600 python bench.py --no-cpu --no-extra --no-e2e --steps 4 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$v', d['value'], d['ms_per_step'], d['roofline_build']['build_ms'])"
done
