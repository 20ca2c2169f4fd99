# A/B of the covariance pass: CTAs per registration (NDTB_COV_CHUNKS), device time of the launches + bench value
for ch in ${CHUNKS:-2 3 4 6 8}; do
  export NDTB_COV_CHUNKS=$ch
  echo "== chunks $ch"
  timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:cov_pass_kernel -s 4 -c 4 --csv python bench.py --steps 1 --warmup 1 --no-cpu --no-e2e --no-extra --lanes 1 2>/dev/null | grep cov_pass | awk -F'","' '{print $NF}' | tr -d '"' | tr '\n' ' '; echo
done
unset NDTB_COV_CHUNKS
python bench.py --no-cpu --no-extra --no-e2e --steps 6 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('value', d['value'], 'ms/step', d['ms_per_step'])"
