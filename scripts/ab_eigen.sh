# A/B of k_eigen launch bounds (scripts/build_variants.sh map_build eigN "-DNDTB_EIGEN_MINBLOCKS=N"): device time of the launches
for v in "" eig5 eig6; do
  if [ -z "$v" ]; then unset NDTB_LIB; else export NDTB_LIB=$PWD/ndt_feature_graph_b200/lib/variants/libndtb_$v.so; fi
  echo "== variant '${v:-default}'"
  timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_eigen -s 8 -c 4 --csv python bench.py --steps 1 --warmup 1 --no-cpu --no-e2e --no-extra --lanes 1 2>/dev/null | grep k_eigen | awk -F'","' '{print $NF}' | tr -d '"' | tr '\n' ' '; echo
done
