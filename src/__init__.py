"""Drop-in mirror of the reference's `src` package: same module paths and public names, implemented by
tarl_simulator_b200 (sm_100a CUDA kernels behind a C ABI). See INTEGRATION.md."""
