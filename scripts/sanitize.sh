#!/bin/bash
# compute-sanitizer over the round-2 kernels (ray trace, radix sort, covariance queue, storage box): memcheck + racecheck
OUT=gpurun_out; mkdir -p $OUT
T="tests/test_gpu_fuser.py::test_ray_trace_bit_exact_vs_oracle tests/test_gpu_fuser.py::test_two_add_point_clouds_before_one_compute tests/test_gpu_fuser.py::test_node7_map_reproduced_exactly_on_gpu tests/test_gpu_parity.py::test_register_scans_matches_stepwise tests/test_gpu_parity.py::test_covariance tests/test_gpu_parity.py::test_load_point_cloud_centroid tests/test_gpu_parity.py::test_overlap_scores_batched tests/test_gpu_parity.py::test_cell_vector_derivatives_and_line_search"
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest $T -q -x > $OUT/sanitizer_memcheck_r02.txt 2>&1; echo "memcheck exit $?" >> $OUT/sanitizer_memcheck_r02.txt
timeout 1500 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_fuser.py::test_two_add_point_clouds_before_one_compute tests/test_gpu_parity.py::test_covariance tests/test_gpu_parity.py::test_register_scans_matches_stepwise -q -x > $OUT/sanitizer_racecheck_r02.txt 2>&1; echo "racecheck exit $?" >> $OUT/sanitizer_racecheck_r02.txt
tail -4 $OUT/sanitizer_memcheck_r02.txt; tail -4 $OUT/sanitizer_racecheck_r02.txt
python scripts/replay_mapping.py --backend gpu --stride 11 --soft 1 --out $OUT/replay_mapping_bag_gpu.json | tail -1
