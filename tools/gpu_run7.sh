set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/tests7.log 2>&1; tail -3 gpurun_out/tests7.log
python tools/probe_nvml.py > gpurun_out/probe_nvml2.log 2>&1; cat gpurun_out/probe_nvml2.log | cut -c1-420
