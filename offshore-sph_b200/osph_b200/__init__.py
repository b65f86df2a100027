"""B200-native WCSPH step: ctypes binding of libosph_b200.so plus host-side helpers."""
