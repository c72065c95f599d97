// empty stand-in for the CUDA header when the device kernels are compiled for the host (tests/emul/emul_kernels.cpp)
