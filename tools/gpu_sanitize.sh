#!/bin/bash
# compute-sanitizer passes over small parity tests (memcheck: out-of-bounds / misaligned; racecheck: shared-memory hazards
# of the per-warp histograms and staging buffers of both kernels)
mkdir -p gpurun_out
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 99 python -m pytest tests -m gpu -x -q -k "narrow or duplicate or fast_xi_wp or weighted_sums or fast_DDtheta_reference or precision_suffixed or (refined_lattice and 6-link0-0-float64) or too_large" > gpurun_out/k_memcheck.log 2>&1
echo "memcheck rc=$?"; grep -E "ERROR SUMMARY|passed|failed" gpurun_out/k_memcheck.log | tail -3
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 99 python -m pytest tests -m gpu -x -q -k "narrow or (fast_xi_wp and float32) or (weighted_sums and signed and float64) or (refined_lattice and 6-link0-1 and float32) or (DDsmu_vs_oracle and 1-True-float64)" > gpurun_out/k_racecheck.log 2>&1
echo "racecheck rc=$?"; grep -E "RACECHECK SUMMARY|passed|failed|hazard" gpurun_out/k_racecheck.log | tail -5
