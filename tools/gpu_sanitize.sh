#!/bin/bash
# compute-sanitizer pass over the CUDA path (the reference has no race / memory checking at all).
# memcheck on the whole small-shape GPU suite in the SIMT GEMM mode plus the SIMT-only kernels (training / evaluation
# tail, RoIAlign, geometry, graph build); racecheck + synccheck on the shared-memory kernels.  The tcgen05 / TMA engines
# are exercised by memcheck only (racecheck does not model the async proxy); keep shapes small: the tools slow kernels
# down 10-100x.   Usage (B200 box):  bash tools/gpu_sanitize.sh [memcheck|racecheck|synccheck|all]
mkdir -p gpurun_out
what=${1:-all}
SAN=/usr/local/cuda/bin/compute-sanitizer
SMALL='tests/test_gpu_train_tail.py::test_losses_vs_golden_and_oracle tests/test_gpu_train_tail.py::test_clip_and_sgd_vs_golden tests/test_gpu_eval_tail.py::test_ties_are_broken_by_edge_id_and_probabilities_pass_through tests/test_gpu_eval_tail.py::test_per_image_ranking_equals_image_by_image tests/test_gpu_parity.py::test_roi_align_node_and_union tests/test_gpu_grad.py::test_linear_backward_vs_torch'
run() {  # tool, extra flags, test selection
  echo "=== $1"; timeout 1500 $SAN --tool $1 $2 --error-exitcode 9 --log-file gpurun_out/sanitize_$1.log \
    python -m pytest $3 -m gpu -x -q -p no:cacheprovider > gpurun_out/sanitize_$1.pytest.log 2>&1
  echo "rc=$? (9 = sanitizer errors)"; tail -3 gpurun_out/sanitize_$1.log; tail -2 gpurun_out/sanitize_$1.pytest.log
}
if [ "$what" = memcheck ] || [ "$what" = all ]; then
  run memcheck "--leak-check no" "$SMALL tests/test_gpu_parity.py::test_l0_message_pass_vs_golden_and_oracle tests/test_gpu_grad.py::test_l1_gradients_vs_reference_autograd"
fi
if [ "$what" = racecheck ] || [ "$what" = all ]; then run racecheck "" "$SMALL"; fi
if [ "$what" = synccheck ] || [ "$what" = all ]; then run synccheck "" "$SMALL"; fi
