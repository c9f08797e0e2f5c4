// stub for tests/cpp/k1_emul.cpp: the host emulation defines the few CUDA names it needs itself
