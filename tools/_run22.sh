mkdir -p gpurun_out
python -m pytest tests/test_gpu_round2.py -q -m gpu -k "conv_relu" 2>&1 | tail -15 | cut -c1-400
python -m pytest tests/test_gpu_models.py tests/test_gpu_dropin.py -q -m gpu -x 2>&1 | tail -5 | cut -c1-300
python -m pytest tests/test_gpu_parity.py -q -m gpu -x -k "fusion or conv2d_oracle or train_trace" 2>&1 | tail -4 | cut -c1-300
