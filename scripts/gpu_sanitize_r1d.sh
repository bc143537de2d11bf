# memcheck / racecheck of what changed in round 1d: the z-slab pipeline (host fabs, skip_top_z),
# velocity sweeps and the per-direction BC flags in k_stream4
timeout 420 compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_pipeline.py "tests/test_gpu_parity.py::test_edge_3d_boundary_box" "tests/test_gpu_parity.py::test_fused_density_pass" -m gpu -x -q 2>&1 | grep -E "passed|failed|ERROR SUMMARY|Invalid|error" | head
timeout 300 compute-sanitizer --tool racecheck python -m pytest "tests/test_gpu_pipeline.py::test_density_pass_host_slabs" "tests/test_gpu_parity.py::test_fused_density_pass" -m gpu -x -q -k "16 or reacting or outflow or mixed" 2>&1 | grep -E "passed|failed|RACECHECK SUMMARY|hazard" | head -5
