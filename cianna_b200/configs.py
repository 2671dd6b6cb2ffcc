"""Network definitions of the BASELINE.json configs, as `spec` dicts for cianna_b200.utils.build_network."""

# "not set" markers of the YOLO helper arrays (src/python_module.c:583-884): error scale -1, fit part -2, IoU limit -2
UNSET_SM = [-1.0, 100000.0, -100000.0]


def darknet19(batch, size=448, classes=1000):
    """The north-star network: Darknet19 ImageNet classifier as defined upstream in
    examples/ImageNET/imagenet_train.py:45-92 (19 convs, 18 group-norms, 5 max-pools, global average pool + softmax)."""
    L = []

    def c(f, n, act="RELU"):
        L.append(("conv", dict(f_size=(f, f), nb_filters=n, padding=(f // 2, f // 2), activation=act)))

    def gn(g):
        L.append(("norm", dict(normalization="GN", group_size=g, set_off=0)))

    def mp():
        L.append(("pool", dict(p_size=(2, 2), p_type="MAX")))

    c(3, 32); gn(4); mp()
    c(3, 64); gn(8); mp()
    c(3, 128); gn(8); c(1, 64); gn(8); c(3, 128); gn(8); mp()
    c(3, 256); gn(16); c(1, 128); gn(16); c(3, 256); gn(16); mp()
    c(3, 512); gn(16); c(1, 256); gn(16); c(3, 512); gn(16); c(1, 256); gn(16); c(3, 512); gn(16); mp()
    c(3, 1024); gn(32); c(1, 512); gn(16); c(3, 1024); gn(32); c(1, 512); gn(16); c(3, 1024); gn(32)
    c(1, classes, "LIN")
    L.append(("pool", dict(p_type="AVG", p_global=1, activation="SMAX")))
    return dict(in_dim=(size, size), in_ch=3, out_dim=classes, bias=0.1, batch=batch, layers=L)


def lenet(batch=16, size=28, d1=256, d2=128, dropout=False):
    """MNIST LeNet-5-like of examples/MNIST/mnist_train.py:67-73 upstream; dropout=True keeps its drop_rate 0.5 / 0.2 on
    the dense layers (off in the deterministic parity fixtures)."""
    drop1 = dict(drop_rate=0.5) if dropout else {}
    drop2 = dict(drop_rate=0.2) if dropout else {}
    return dict(in_dim=(size, size), in_ch=1, out_dim=10, bias=0.1, batch=batch, layers=[
        ("conv", dict(f_size=(5, 5), nb_filters=8, padding=(2, 2), activation="RELU")),
        ("pool", dict(p_size=(2, 2), p_type="MAX")),
        ("conv", dict(f_size=(5, 5), nb_filters=16, padding=(2, 2), activation="RELU")),
        ("pool", dict(p_size=(2, 2), p_type="MAX")),
        ("dense", dict(nb_neurons=d1, strict_size=1, activation="RELU", **drop1)),
        ("dense", dict(nb_neurons=d2, strict_size=1, activation="RELU", **drop2)),
        ("dense", dict(nb_neurons=10, strict_size=1, activation="SMAX")),
    ])


def darknet19_yolo(batch, size=416, nb_class=80, nb_box=5, max_nb_obj=70):
    """BASELINE config 3: the COCO detector of examples/COCO/coco_train.py:23-90 upstream - the first 40 layers of the
    Darknet19 classifier (up to the last 3x3x1024 conv), then GN + 3 x (3x3x1024 conv + GN) and a 1x1 YOLO head on the
    size/32 grid; YOLO set-up of that script (5 priors, 80 soft-max classes, "difficult" flag, DIoU, strict association)."""
    back = darknet19(batch, size)["layers"][:40]
    L = list(back)
    L.append(("norm", dict(normalization="GN", group_size=16, set_off=0)))
    for _ in range(3):
        L.append(("conv", dict(f_size=(3, 3), nb_filters=1024, padding=(1, 1), activation="RELU")))
        L.append(("norm", dict(normalization="GN", group_size=16, set_off=0)))
    L.append(("conv", dict(f_size=(1, 1), nb_filters=nb_box * (8 + nb_class), padding=(0, 0), activation="YOLO")))
    s = size / 416.0
    y = dict(nb_box=nb_box, nb_class=nb_class, max_nb_obj_per_image=max_nb_obj,
             prior_size=[[32.0 * s, 92.0 * s, 150.0 * s, 208.0 * s, 333.0 * s][:nb_box], [32.0 * s, 150.0 * s, 92.0 * s, 333.0 * s, 208.0 * s][:nb_box]],
             prior_noobj_prob=[0.05, 0.1, 0.1, 0.1, 0.1][:nb_box], IoU_type="DIoU", prior_dist_type="OFFSET",
             error_scales=[12.0, 6.0, 0.5, 6.0, 0.4, -1.0],
             slopes_and_maxes=[[1.0, 6.0, -6.0], [0.5, 1.6, -1.6], [1.0, 6.0, -6.0], [1.0, 6.0, -6.0], [1.0, 6.0, -6.0], UNSET_SM],
             IoU_limits=[0.5, -0.1, -1.0, -1.0, -0.1, -1.0, 0.5, 0.3], fit_parts=[1, 1, 1, 1, 1, -2],
             strict_box_size=3, min_prior_forced_scaling=1.2, diff_flag=1, rand_startup=0, rand_prob_best_box_assoc=0.05,
             class_softmax=1, error_type="natural", no_override=1)
    return dict(in_dim=(size, size), in_ch=3, out_dim=1 + max_nb_obj * (7 + 1), bias=0.1, batch=batch, yolo=y, layers=L)


def sdc1_yolo(batch, size=512, nb_box=9, nb_param=5):
    """BASELINE config 4: the SKA SDC1 source detector of examples/SKAO_SDC1/train_network.py:36-137 upstream (17 conv
    layers, stride-2 2x2 convolutions instead of pooling, one group-norm, dropout 0.25 before the YOLO head, 9 priors, no
    classes, 5 extra parameters per box); 340 target slots per 256 x 256 px as in aux_fct.py:90."""
    L = []

    def c(f, n, s=1, act="RELU", **kw):
        L.append(("conv", dict(f_size=(f, f), nb_filters=n, stride=(s, s), padding=((f // 2) if s == 1 else 0,) * 2, activation=act, **kw)))

    c(5, 32); c(2, 16, 2); c(3, 24); c(3, 32); c(2, 64, 2); c(1, 128); c(3, 192); c(2, 128, 2); c(1, 192); c(3, 384)
    c(1, 256); c(3, 384); c(2, 512, 2); c(1, 768); c(3, 1024)
    L.append(("norm", dict(normalization="GN", group_size=4, set_off=0)))
    c(1, 2048, drop_rate=0.25)
    c(1, nb_box * (8 + nb_param), act="YOLO")
    max_nb_obj = int(340 * (size * size) / (256 * 256))
    y = dict(nb_box=nb_box, nb_class=0, nb_param=nb_param, max_nb_obj_per_image=max_nb_obj,
             prior_size=[[6.0] * 6 + [12.0, 9.0, 24.0], [6.0] * 6 + [9.0, 12.0, 24.0]],
             prior_noobj_prob=[0.15] * 6 + [0.01] * 3, IoU_type="DIoU", prior_dist_type="OFFSET",
             error_scales=[36.0, 0.2, 0.5, 2.0, -1.0, 5.0], param_ind_scales=[2.0, 2.0, 1.0, 0.5, 0.5],
             slopes_and_maxes=[[0.5, 6.0, -6.0], [0.5, 1.2, -1.2], [0.2, 6.0, -6.0], [0.5, 6.0, -6.0], UNSET_SM, [0.5, 1.5, -0.2]],
             IoU_limits=[0.5, -0.1, -0.3, -0.3, -2.0, -0.1, -2.0, -2.0], fit_parts=[1, 1, 1, 1, -2, 1],
             strict_box_size=0, min_prior_forced_scaling=0.0, rand_startup=0, rand_prob_best_box_assoc=0.90, rand_prob=0.02,
             error_type="natural", no_override=1)
    return dict(in_dim=(size, size), in_ch=1, out_dim=1 + max_nb_obj * (7 + nb_param), bias=0.1, batch=batch, yolo=y, layers=L)


def extinction_profile(batch, size=64, in_ch=1, out_dim=128):
    """BASELINE config 5: the dense-GEMM-heavy regression network [C5x5.12 - P2 - D3072 x 2 - D2048 - D128] on 64 px maps
    (quadratic loss on a 128-bin profile, linear output)."""
    return dict(in_dim=(size, size), in_ch=in_ch, out_dim=out_dim, bias=0.1, batch=batch, layers=[
        ("conv", dict(f_size=(5, 5), nb_filters=12, padding=(2, 2), activation="RELU")),
        ("pool", dict(p_size=(2, 2), p_type="MAX")),
        ("dense", dict(nb_neurons=3072, strict_size=1, activation="RELU", drop_rate=0.1)),
        ("dense", dict(nb_neurons=3072, strict_size=1, activation="RELU", drop_rate=0.1)),
        ("dense", dict(nb_neurons=2048, strict_size=1, activation="RELU")),
        ("dense", dict(nb_neurons=out_dim, strict_size=1, activation="LIN")),
    ])
