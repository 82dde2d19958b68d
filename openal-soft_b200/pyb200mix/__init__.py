"""Python-side plumbing for the b200mix C ABI (ctypes mirrors + synthetic scenes).
The product is openal-soft_b200/libb200mix.so; nothing here mixes audio."""
