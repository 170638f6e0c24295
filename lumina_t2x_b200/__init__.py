"""B200-native Next-DiT denoising engine (drop-in for the Lumina-T2X sampling hot path).

``lumina_t2x_b200.models`` / ``lumina_t2x_b200.transport`` mirror the reference's
``models`` / ``transport`` modules; both call libndit_b200.so (include/ndit.h) through ctypes."""
__all__ = ["models", "transport"]
