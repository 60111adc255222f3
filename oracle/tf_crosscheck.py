"""Pin the oracle against the REAL reference (run this wherever TensorFlow 2.7 + tensorflow-addons 0.15 exist).

TEST INFRASTRUCTURE.  The build image has neither package (no wheels, no network), which is why every
parity claim in this repo is "vs. the restated oracle" (PARITY UNPINNED).  This
script is the missing measurement: it imports the reference's own `Superresolution` / `Optimizer`
from a checkout of nicoloalbergoni/DeepLabV3Plus-Augmented-SuperResolution, feeds it the same synthetic
inputs the goldens use, and reports oracle-vs-TensorFlow max-abs per stage:

    forward residual, gradient at x0, x after 1 / 10 / num_iter steps, thresholded-mask agreement.

Usage:
    python oracle/tf_crosscheck.py --reference /path/to/DeepLabV3Plus-Augmented-SuperResolution [--device cpu]
                                   [--write-golden tests/golden/tf_crosscheck.npz] [--write-hdf5 tests/golden/h5py_fixture.hdf5]

--write-golden stores the TensorFlow-produced vectors (inputs, transform matrices and inverses from tfa's helpers, loss, gradient,
x after 1 / 10 / num_iter steps); tests/test_oracle.py::test_tf_generated_goldens then pins the oracle to them on every run.
--write-hdf5 writes one augmented-copies file with real h5py exactly as augmentation_utils.py:123-136 does;
tests/test_host.py::test_hdf5_reader_on_h5py_fixture then pins hdf5_lite's reader to it.  Neither file exists in this repo yet:
the attempt to install TensorFlow / h5py on the build and GPU images is logged in profiles/r02_tf_install_attempt.log.

Expected outcome if the operator semantics of SURVEY.md Appendix A are right: residual and gradient agree to
fp32 rounding (<= ~1e-5 relative); x after many steps agrees to the fp32 "chaos floor" measured in
profiles/r01_chaos_floor.txt (max-abs ~1e-3 at a handful of pixels, mean ~2e-6, masks identical), because
TensorFlow's own CPU and GPU builds already differ from each other by that much.
"""
import argparse
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reference", required=True, help="checkout of the reference repository")
    ap.add_argument("--device", default="cpu", choices=["cpu", "gpu"])
    ap.add_argument("--num-aug", type=int, default=20)
    ap.add_argument("--lr-size", type=int, default=64)
    ap.add_argument("--iters", type=int, default=60)
    ap.add_argument("--write-golden", default=None)
    ap.add_argument("--write-hdf5", default=None)
    args = ap.parse_args()
    if args.device == "cpu":
        os.environ["CUDA_VISIBLE_DEVICES"] = ""
    sys.path.insert(0, args.reference)
    import tensorflow as tf  # noqa: F401  (fails loudly where TF is absent: that is the point)
    from superresolution_scripts.optimizer import Optimizer as RefOptimizer
    from superresolution_scripts.superresolution import Superresolution as RefSR

    from oracle import oracle as O
    from deeplabv3plus_augmented_superresolution_b200.synthetic import make_augmented_copies

    h = args.lr_size
    H = 4 * h
    copies, ang, sh = make_augmented_copies(1, args.num_aug, (h, h), (H, H), 0.15, 80 * H / 512, seed=1234)
    c = copies[0].numpy()
    a, s = ang[0], sh[0]
    kw = dict(lambda_df=1.0, lambda_tv=0.3, lambda_l2=0.7, lambda_l1=0.0)

    def ref_solver(n_iter):
        opt = RefOptimizer(optimizer="adam", learning_rate=1e-3, amsgrad=True, lr_scheduler=True, decay_steps=60, decay_rate=0.3)
        return RefSR(lambda_df=1.0, lambda_tv=0.3, lambda_L2=0.7, lambda_L1=0.0, num_iter=n_iter, num_aug=args.num_aug,
                     optimizer=opt, feature_size=(h, h), output_size=(H, H))

    # stage 1: loss and gradient at x0
    sr = ref_solver(1)
    x0 = tf.image.resize(c[0][..., None], (H, H))[tf.newaxis]
    xv = tf.Variable(x0)
    with tf.GradientTape() as tape:
        loss = sr.loss_function(xv, tf.constant(c[..., None]), tf.constant(a), tf.constant(s), n_drop=0)
    g_tf = tape.gradient(loss, [xv])[0].numpy()[0, :, :, 0]
    lo, g_o = O.loss_and_grad(x0.numpy()[0, :, :, 0], c, a, s, O.SolveParams(**kw))
    print(f"loss        tf {float(loss):.6f}  oracle {lo:.6f}")
    print(f"gradient    max-abs {np.abs(g_tf - g_o).max():.3e}  rel-L2 {np.linalg.norm(g_tf - g_o) / np.linalg.norm(g_tf):.3e}")

    gold = {"copies": c, "angles": a, "shifts": s, "H": H, "loss_tf": np.float32(loss), "grad_tf": g_tf, "iters": np.int32(args.iters)}
    try:   # the transform helpers the reference reaches through tfa.image.rotate / translate, and tf.linalg.inv of the 3x3
        import tensorflow_addons as tfa
        rot = tfa.image.angles_to_projective_transforms(tf.constant(a), float(H), float(H)).numpy()
        tr = tfa.image.translations_to_projective_transforms(tf.constant(s)).numpy()
        def inv8(t):
            m = tf.reshape(tf.concat([t, tf.ones((t.shape[0], 1), tf.float32)], 1), (-1, 3, 3))
            mi = tf.linalg.inv(m)
            return (tf.reshape(mi, (-1, 9)) / mi[:, 2, 2, None])[:, :8].numpy()
        gold.update(rot_tf=rot, tr_tf=tr, rot_inv_tf=inv8(tf.constant(rot)), tr_inv_tf=inv8(tf.constant(tr)))
        print(f"rotate matrices  max-abs vs oracle {max(np.abs(rot[k] - O.rotate_matrix(a[k], H, H)).max() for k in range(len(a))):.3e}")
    except Exception as e:   # tfa may be missing where core TF exists
        print("tensorflow_addons helpers unavailable:", e)

    # stage 2: iterates
    for n in (1, 10, args.iters):
        x_tf, _ = ref_solver(n).augmented_superresolution(tf.constant(c[..., None]), a, s)
        x_o, _ = O.augmented_superresolution(c, a, s, O.SolveParams(num_iter=n, **kw), output_size=(H, H))
        d = np.abs(x_tf - x_o)
        m_tf = O.threshold_image(x_tf, 8, th_factor=0.65)
        m_o = O.threshold_image(x_o, 8, th_factor=0.65)
        print(f"x after {n:4d}  max-abs {d.max():.3e}  mean-abs {d.mean():.3e}  pixels>1e-4 {(d > 1e-4).sum()}  "
              f"mask agreement {(m_tf == m_o).mean():.6f}")
        gold[f"x_tf_{n}"] = np.asarray(x_tf, np.float32)
    if args.write_golden:
        np.savez_compressed(args.write_golden, **gold)
        print("wrote", args.write_golden)
    if args.write_hdf5:
        import h5py
        f = h5py.File(args.write_hdf5, "w")                      # augmentation_utils.py:123-136, verbatim calls
        f.create_dataset("class_masks", data=[m[..., None] for m in c])
        f.create_dataset("angles", data=a)
        f.create_dataset("shifts", data=s)
        f.attrs["filename"] = "2007_000032"
        f.attrs["mode"] = "argmax"
        f.attrs["angle_max"] = 0.15
        f.attrs["shift_max"] = 80
        f.close()
        np.save(args.write_hdf5 + ".class_masks.npy", np.stack([m[..., None] for m in c]))
        print("wrote", args.write_hdf5)


if __name__ == "__main__":
    main()
