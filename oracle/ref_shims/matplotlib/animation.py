class ImageMagickWriter:
    def __init__(self, *a, **k):
        pass
