"""Per-dataset CaRA hyper-parameters (values of the reference's image_classification/vtab_config.py).

``config[name] = {init_mean, init_std, scale, seed, logger}``: ``scale`` is the adapter scale ``s`` of the
fused kernels, ``init_mean``/``init_std`` initialise ``CP_R1``/``CP_R2`` (vit_cp.py:131-134,156-162).
"""

#          dataset                init_mean init_std scale  seed
_TABLE = (("cifar",                 1.5,   0.1,    0.1,   14),
          ("caltech101",            0.9,   0.01,   100,   56),
          ("dtd",                   1.0,   0.0,    0.1,   14),
          ("oxford_flowers102",     1.0,   0.02,   10.0,  50),
          ("oxford_iiit_pet",       1.2,   0.06,   1.0,   93),
          ("svhn",                  1.0,   0.05,   100,   14),
          ("sun397",                1.35,  0.06,   1.0,   43),
          ("patch_camelyon",        1.0,   0.0,    10,    89),
          ("eurosat",               1.08,  0.028,  10,    32),
          ("resisc45",              1.16,  0.03,   10,    28),
          ("diabetic_retinopathy",  1.0,   0.0,    0.1,   81),
          ("clevr_count",           1.0,   0.0,    5,     44),
          ("clevr_dist",            1.0,   0.0,    2.5,   25),
          ("dmlab",                 1.0,   0.0,    10,    72),
          ("kitti",                 1.0,   0.0,    5,     31),
          ("dsprites_loc",          1.0,   0.0,    50,    12),
          ("dsprites_ori",          1.3,   0.07,   1.0,   79),
          ("smallnorb_azi",         1.0,   0.0,    100,   67),
          ("smallnorb_ele",         1.0,   0.0,    10.0,  30))

config = {name: {"init_mean": mu, "init_std": sd, "scale": sc, "seed": seed, "logger": False}
          for name, mu, sd, sc, seed in _TABLE}
