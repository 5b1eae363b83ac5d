def _unavailable(*a, **k):
    raise RuntimeError("skimage is not installed in this image; this augmentation path is outside the hot path")


gaussian = denoise_bilateral = find_boundaries = disk = dilation = _unavailable
