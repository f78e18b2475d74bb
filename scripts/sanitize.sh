# compute-sanitizer passes over small instances of every kernel (memcheck: out-of-bounds / misaligned; racecheck: shared-memory hazards)
set -x
S="compute-sanitizer --error-exitcode 7 --print-limit 5"
$S --tool memcheck python -m pytest tests/test_gpu_fmpc_gen.py -m gpu -x -q -k "s301 or s304 or s306 or s308 or s310 or cold_start" 2>&1 | tail -4
$S --tool memcheck python -m pytest tests/test_gpu_fmpc.py -m gpu -x -q -k "s1_ or s7_ or s9_ or s14_ or s21_ or s23_ or s8_" 2>&1 | tail -4
$S --tool memcheck python -m pytest tests/test_gpu_zernike.py tests/test_gpu_estimator.py tests/test_gpu_varid.py -m gpu -x -q -k "16-2 or 33-0 or 20-11 or 50-4 or 64-2 or other_shapes or 6-100 or 1-50 or synthesis_matches_basis" 2>&1 | tail -4
$S --tool racecheck python -m pytest tests/test_gpu_fmpc.py -m gpu -x -q -k "s23_ or s21_" 2>&1 | tail -4
$S --tool racecheck python -m pytest tests/test_gpu_fmpc_gen.py -m gpu -x -q -k "s310 or s305" 2>&1 | tail -4
$S --tool racecheck python -m pytest tests/test_gpu_zernike.py tests/test_gpu_varid.py -m gpu -x -q -k "16-2 or 64-2-130 or 16-12-3 or 1-50" 2>&1 | tail -4
# round 2: resident step (fused shifted warm start / U(:,0) extraction and the separate shift / extract kernels), the device MT19937
# generator (memcheck + racecheck: one CTA, double-buffered blocks), the multi-device entry points, dense R, arbitrary sample sets
$S --tool memcheck python -m pytest tests/test_gpu_fmpc.py tests/test_gpu_multi.py -m gpu -x -q -k "resident or stream or multi" 2>&1 | tail -4
$S --tool racecheck python -m pytest tests/test_gpu_fmpc.py -m gpu -x -q -k "stream_on_device or resident_steps" 2>&1 | tail -4
$S --tool memcheck python -m pytest tests/test_gpu_fmpc_gen.py tests/test_gpu_zernike.py tests/test_gpu_estimator.py -m gpu -x -q -k "s401 or s403 or s404 or 500-4 or 257-10 or other_shapes" 2>&1 | tail -4
$S --tool racecheck python -m pytest tests/test_gpu_fmpc_gen.py -m gpu -x -q -k "s401 or s403" 2>&1 | tail -4
# general-structure kernel after the round-2 rewrite: both block sizes, n > 32 (shared-memory potrf), dense R, chunked panel (n = 100)
$S --tool memcheck python -m pytest tests/test_gpu_fmpc_gen.py -m gpu -x -q -k "s301 or s304 or s306 or s308 or s309 or s310 or cold_start or s401 or s403" 2>&1 | tail -4
$S --tool racecheck python -m pytest tests/test_gpu_fmpc_gen.py -m gpu -x -q -k "s310 or s305 or s304 or s401" 2>&1 | tail -4
$S --tool memcheck python -m pytest tests/test_gpu_fmpc.py -m gpu -x -q -k "large_state_dimension" 2>&1 | tail -4
