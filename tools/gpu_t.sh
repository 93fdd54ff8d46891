timeout 1500 python -m pytest tests/test_chamfer_gpu.py -m gpu -q -x > gpurun_out/t_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/t_pytest.log
