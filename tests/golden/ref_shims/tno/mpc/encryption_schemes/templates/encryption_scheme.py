class EncodedPlaintext:
    def __init__(self, value, scheme=None):
        self.value = value
        self.scheme = scheme

    def __class_getitem__(cls, _item):
        return cls
