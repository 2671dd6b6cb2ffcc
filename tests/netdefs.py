"""Network specs shared by the reference side (oracle/ref_driver.py) and the product side."""


def mini_darknet(batch=4, size=16, classes=6, in_ch=3, width=8):
    """Darknet19-shaped slice (examples/ImageNET/imagenet_train.py:45-92 upstream): C3x3-GN-MaxP, C3x3-GN-C1x1-GN-C3x3-GN-MaxP,
    C1x1 LIN head, global average pool + softmax."""
    w = width
    return dict(in_dim=(size, size), in_ch=in_ch, out_dim=classes, bias=0.1, batch=batch, layers=[
        ("conv", dict(f_size=(3, 3), nb_filters=w, padding=(1, 1), activation="RELU")),
        ("norm", dict(normalization="GN", group_size=4)),
        ("pool", dict(p_size=(2, 2), p_type="MAX")),
        ("conv", dict(f_size=(3, 3), nb_filters=2 * w, padding=(1, 1), activation="RELU")),
        ("norm", dict(normalization="GN", group_size=8)),
        ("conv", dict(f_size=(1, 1), nb_filters=w, padding=(0, 0), activation="RELU")),
        ("norm", dict(normalization="GN", group_size=4)),
        ("conv", dict(f_size=(3, 3), nb_filters=2 * w, padding=(1, 1), activation="RELU")),
        ("norm", dict(normalization="GN", group_size=8)),
        ("pool", dict(p_size=(2, 2), p_type="MAX")),
        ("conv", dict(f_size=(1, 1), nb_filters=classes, padding=(0, 0), activation="LIN")),
        ("pool", dict(p_type="AVG", p_global=1, activation="SMAX")),
    ])


def tc_darknet(batch=8, size=16, classes=16):
    """Same shape with channel counts that route every conv but the first through the tcgen05 kernels
    (>= 16 input channels, multiples of 16)."""
    return dict(in_dim=(size, size), in_ch=3, out_dim=classes, bias=0.1, batch=batch, layers=[
        ("conv", dict(f_size=(3, 3), nb_filters=32, padding=(1, 1), activation="RELU")),
        ("norm", dict(normalization="GN", group_size=4)),
        ("pool", dict(p_size=(2, 2), p_type="MAX")),
        ("conv", dict(f_size=(3, 3), nb_filters=64, padding=(1, 1), activation="RELU")),
        ("norm", dict(normalization="GN", group_size=8)),
        ("conv", dict(f_size=(1, 1), nb_filters=32, padding=(0, 0), activation="RELU")),
        ("norm", dict(normalization="GN", group_size=8)),
        ("conv", dict(f_size=(3, 3), nb_filters=128, padding=(1, 1), activation="RELU")),
        ("norm", dict(normalization="GN", group_size=16)),
        ("pool", dict(p_size=(2, 2), p_type="MAX")),
        ("conv", dict(f_size=(1, 1), nb_filters=classes, padding=(0, 0), activation="LIN")),
        ("pool", dict(p_type="AVG", p_global=1, activation="SMAX")),
    ])


from cianna_b200.configs import darknet19, lenet  # noqa: E402,F401  (shared with bench.py)
