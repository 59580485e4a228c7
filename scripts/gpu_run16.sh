set -x
timeout 300 python -m pytest tests -m gpu -x -q -k "peer_memory" 2>&1 | tail -8
