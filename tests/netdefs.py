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


def dropout_net(batch=6, size=12):
    """dropout on every layer kind that has it upstream (conv: examples/SKAO_SDC1/train_network.py:136, dense:
    examples/MNIST/mnist_train.py:71-72, pool), rates on and off 1/65536 steps"""
    return dict(in_dim=(size, size), in_ch=2, out_dim=5, bias=0.1, batch=batch, layers=[
        ("conv", dict(f_size=(3, 3), nb_filters=12, padding=(1, 1), activation="RELU", drop_rate=0.25)),
        ("conv", dict(f_size=(3, 3), nb_filters=16, padding=(1, 1), activation="RELU")),
        ("pool", dict(p_size=(2, 2), p_type="MAX", drop_rate=0.3)),
        ("conv", dict(f_size=(1, 1), nb_filters=8, activation="LIN", drop_rate=0.1)),
        ("dense", dict(nb_neurons=24, strict_size=1, activation="RELU", drop_rate=0.5)),
        ("dense", dict(nb_neurons=12, strict_size=1, activation="RELU", drop_rate=0.2)),
        ("dense", dict(nb_neurons=5, strict_size=1, activation="SMAX")),
    ])


def regression_net(out_act="LIN", head="dense", batch=5, size=10):
    """quadratic-loss networks (the regression use of upstream, e.g. its 3-D extinction-profile model): logistic hidden
    layers and a LIN / RELU / LOGI output layer, dense or convolutional"""
    layers = [
        ("conv", dict(f_size=(3, 3), nb_filters=8, padding=(1, 1), activation="LOGI")),
        ("pool", dict(p_size=(2, 2), p_type="AVG")),
    ]
    if head == "dense":
        layers += [("dense", dict(nb_neurons=16, strict_size=1, activation="LOGI")),
                   ("dense", dict(nb_neurons=6, strict_size=1, activation=out_act))]
        out_dim = 6
    else:
        layers += [("conv", dict(f_size=(3, 3), nb_filters=4, padding=(1, 1), activation=out_act))]
        out_dim = 4 * (size // 2) ** 2
    return dict(in_dim=(size, size), in_ch=2, out_dim=out_dim, bias=0.1, batch=batch, layers=layers)


REGRESSION_CASES = [("LIN", "dense"), ("RELU", "dense"), ("LOGI", "dense"), ("RELU", "conv"), ("LOGI", "conv")]


def yolo_head(name, batch=3, grid=6, cell=8):
    """One-layer network whose only layer is the YOLO head (conv with filter = stride = cell): the kernel-level parity
    cases of tests/test_gpu_yolo.py and of the golden fixtures tests/golden/yolo_<name>.npz.  Every case forces the
    deterministic association branches (rand_startup=0, no random re-association)."""
    det = dict(rand_startup=0, rand_prob=0.0, rand_prob_best_box_assoc=0.0)
    cases = {
        # upstream defaults: GIoU, closest prior by size, no strict association
        "giou_default": dict(nb_box=3, nb_class=2, nb_param=2, max_nb_obj_per_image=6, prior_size=[[6., 12., 20.], [6., 10., 22.]],
                             IoU_type="GIoU", prior_dist_type="SIZE", error_type="complete", **det),
        # classical IoU, strict association to the 2 closest priors (one prior duplicated), prior distance by IoU
        "iou_strict": dict(nb_box=5, nb_class=3, nb_param=0, max_nb_obj_per_image=8,
                           prior_size=[[5., 9., 9., 16., 24.], [5., 9., 9., 14., 26.]], IoU_type="IoU", prior_dist_type="IoU",
                           strict_box_size=2, error_type="complete", **det),
        # DIoU, prior distance on the log size offsets, softmax classes, 1 allowed prior
        "diou_offset_softmax": dict(nb_box=4, nb_class=4, nb_param=1, max_nb_obj_per_image=5,
                                    prior_size=[[4., 8., 14., 22.], [4., 9., 12., 24.]], IoU_type="DIoU", prior_dist_type="OFFSET",
                                    strict_box_size=1, class_softmax=1, error_type="complete", **det),
        # DIoU2 with "difficult" flags and the natural error display
        "diou2_difficult": dict(nb_box=3, nb_class=2, nb_param=1, max_nb_obj_per_image=6, prior_size=[[6., 12., 20.], [6., 10., 22.]],
                                IoU_type="DIoU2", prior_dist_type="SIZE", diff_flag=1, error_type="natural", **det),
        # user tables: partial fits, custom scales / slopes / limits (a high low-IoU limit sends most targets to their best prior)
        "custom_tables": dict(nb_box=3, nb_class=2, nb_param=2, max_nb_obj_per_image=6, prior_size=[[6., 12., 20.], [6., 10., 22.]],
                              IoU_type="GIoU", prior_dist_type="SIZE", error_type="complete",
                              fit_parts=[1, 0, 1, 1, 0, 0], error_scales=[1.5, 0.5, 2.0, 1.0, 0.7, 3.0],
                              slopes_and_maxes=[[1.2, 5.0, -5.0], [0.8, 1.2, -1.4], [1.0, 6.0, -6.0], [1.5, 4.0, -4.0], [1.0, 6.0, -6.0], [0.5, 1.5, -0.5]],
                              IoU_limits=[0.3, 0.25, -0.2, -0.1, 0.0, 0.1, 0.4, 0.2], prior_noobj_prob=[0.1, 0.3, 0.5],
                              param_ind_scales=[2.0, 0.5], **det),
        # no classes / params, one box per cell, more targets than boxes in crowded cells
        "single_box": dict(nb_box=1, nb_class=0, nb_param=0, max_nb_obj_per_image=10, prior_size=[[10.], [10.]],
                           IoU_type="GIoU", prior_dist_type="SIZE", error_type="complete", **det),
    }
    y = cases[name]
    per = 7 + y.get("nb_param", 0) + y.get("diff_flag", 0)
    nf = y["nb_box"] * (8 + y.get("nb_class", 0) + y.get("nb_param", 0))
    return dict(in_dim=(grid * cell, grid * cell), in_ch=1, out_dim=1 + y["max_nb_obj_per_image"] * per, bias=0.1, batch=batch, yolo=y,
                layers=[("conv", dict(f_size=(cell, cell), stride=(cell, cell), nb_filters=nf, activation="YOLO"))])


YOLO_HEAD_CASES = ("giou_default", "iou_strict", "diou_offset_softmax", "diou2_difficult", "custom_tables", "single_box")


def yolo_net(batch=4, size=32):
    """small detector: two conv+pool stages and a 1x1 YOLO head on an 8x8 grid (cell = 4 px)"""
    y = dict(nb_box=3, nb_class=2, nb_param=1, max_nb_obj_per_image=6, prior_size=[[5., 9., 14.], [5., 8., 16.]],
             IoU_type="GIoU", prior_dist_type="SIZE", error_type="complete", rand_startup=0)
    return dict(in_dim=(size, size), in_ch=3, out_dim=1 + 6 * 8, bias=0.1, batch=batch, yolo=y, layers=[
        ("conv", dict(f_size=(3, 3), nb_filters=16, padding=(1, 1), activation="RELU")),
        ("pool", dict(p_size=(2, 2), p_type="MAX")),
        ("conv", dict(f_size=(3, 3), nb_filters=32, padding=(1, 1), activation="RELU")),
        ("norm", dict(normalization="GN", group_size=8)),
        ("pool", dict(p_size=(2, 2), p_type="MAX")),
        ("conv", dict(f_size=(1, 1), nb_filters=3 * (8 + 2 + 1), activation="YOLO")),
    ])


from cianna_b200.configs import darknet19, lenet  # noqa: E402,F401  (shared with bench.py)


def lrn_net(batch=4):
    """two LRN layers (one with explicit parameters, one with upstream's defaults) between conv / pool / dense layers"""
    return dict(in_dim=(12, 12), in_ch=3, out_dim=5, bias=0.1, batch=batch, layers=[
        ("conv", dict(f_size=(3, 3), nb_filters=12, padding=(1, 1), activation="RELU")),
        ("lrn", dict(range=5, k=2.0, alpha=0.3, beta=0.75)),
        ("pool", dict(p_size=(2, 2), p_type="MAX")),
        ("conv", dict(f_size=(3, 3), nb_filters=16, padding=(1, 1), activation="RELU")),
        ("lrn", dict()),
        ("dense", dict(nb_neurons=5, strict_size=1, activation="SMAX")),
    ])
