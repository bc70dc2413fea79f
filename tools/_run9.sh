timeout 100 python tools/stats_icp_sofa.py 2.0 100 0.7 2>&1 | tail -1
timeout 100 python tools/stats_icp_sofa.py 2.0 100 1.0 2>&1 | tail -1
timeout 900 python -m pytest tests/test_gpu_icp.py tests/test_gpu_pipeline.py tests/test_gpu_full_size.py -x -q 2>&1 | tail -3
