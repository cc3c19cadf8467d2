"""B200-native DepthFirstRenderer path of gsm-renderer (host-side mirror + CUDA C-ABI)."""
