#!/bin/bash
# Runs the -m gpu parity tests in separate processes (a sticky CUDA error in one group cannot poison
# the others) and writes logs under gpurun_out/. Usage: scripts/gpu_tests.sh [extra pytest args]
mkdir -p gpurun_out
run() {
    local name=$1; shift
    timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q --timeout 300 --timeout-method=thread \
        -p no:cacheprovider --tb=short "$@" > gpurun_out/pytest_${name}.log 2>&1
    echo "[$name] rc=$? $(tail -1 gpurun_out/pytest_${name}.log)"
}
run golden_f32 -k "golden_fixture and float32"
run golden_f64 -k "golden_fixture and float64"
run oracle_f32 -k "stream_equals_oracle and float32"
run oracle_f64 -k "stream_equals_oracle and float64"
run misc -k "not golden_fixture and not stream_equals_oracle and not baseline_config and not large_configs"
run baseline -k "baseline_config"
run large -k "large_configs"
timeout 600 python -m pytest tests/test_cpp_adapter.py -m gpu -q -p no:cacheprovider > gpurun_out/pytest_adapter.log 2>&1; echo "[adapter] rc=$? $(tail -1 gpurun_out/pytest_adapter.log)"
