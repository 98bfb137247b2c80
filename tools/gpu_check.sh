python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/s2_pytest.log
python __graft_entry__.py --smoke > gpurun_out/s2_smoke.log 2>&1
python tools/h2d_bw.py > gpurun_out/s2_h2d.log 2>&1
python bench.py > gpurun_out/s2_bench.json 2> gpurun_out/s2_bench.err
nvidia-smi -q | grep -i -A3 "pci" | head -40 > gpurun_out/s2_pci.log
cat gpurun_out/s2_pytest.log gpurun_out/s2_smoke.log gpurun_out/s2_h2d.log
