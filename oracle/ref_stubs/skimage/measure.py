def compare_ssim(*a, **k):
    raise NotImplementedError("stub")
