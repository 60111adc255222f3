"""Drop-in for superresolution_scripts/superresolution.py (reference :26-161) on libasr.

Same constructor kwargs, same three methods, same return types (NumPy [H,W,1] float32 plus the
last loss, or None for the max/mean baselines).  The TensorFlow op graph of loss_function /
tape.gradient / apply_gradients (reference :44-137) is replaced by asr_solve_batched; the extra
`*_batched` methods expose what the C ABI adds: many images (or many hyper-parameter points) per
call, device-resident inputs and outputs.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from .. import _lib
from .optimizer import Optimizer


def _as_device_stack(copies):
    """Accepts what the reference's callers pass (a list of [h,w,1] arrays, an [N,h,w,1] array or a
    tensor; test_SR.py:73-77, superres_utils.py:178-194) and returns a contiguous CUDA [N,h,w] float32."""
    torch = _lib._torch()
    if isinstance(copies, torch.Tensor):
        t = copies
    else:
        if isinstance(copies, (list, tuple)):
            copies = np.stack([c.detach().cpu().numpy() if isinstance(c, torch.Tensor) else np.asarray(c)
                               for c in copies])
        t = torch.from_numpy(np.ascontiguousarray(copies, dtype=np.float32))
    if t.dim() == 4 and t.shape[-1] == 1:
        t = t[..., 0]
    if t.dim() != 3:
        raise ValueError(f"expected [N,h,w,1] or [N,h,w] copies, got {tuple(t.shape)}")
    return t.to(device="cuda", dtype=torch.float32).contiguous()


class Superresolution:
    def __init__(self, lambda_df, lambda_tv, lambda_L2, lambda_L1, num_iter=200, num_aug=100, optimizer: Optimizer = None,
                 feature_size=(64, 64), output_size=(512, 512), use_BTV=False, verbose=False, copy_dropout=0.0):
        self.lambda_df = lambda_df
        self.lambda_tv = lambda_tv
        self.lambda_L2 = lambda_L2
        self.lambda_L1 = lambda_L1

        self.num_iter = num_iter
        self.num_aug = num_aug
        self.optimizer = optimizer
        self.feature_size = feature_size
        self.output_size = output_size
        self.use_BTV = use_BTV
        self.verbose = verbose
        self.copy_dropout = copy_dropout
        self._keep_mask = None   # drawn once, like the mask frozen at tf.function trace time (reference :47-53)

    # ---- parameters ------------------------------------------------------------------------------------------
    def _solve_params(self, step_offset: int, **overrides) -> _lib.SolveParams:
        o = self.optimizer
        p = _lib.SolveParams(
            lambda_df=float(self.lambda_df), lambda_tv=float(self.lambda_tv), lambda_l2=float(self.lambda_L2),
            lambda_l1=float(self.lambda_L1), num_iter=int(self.num_iter), optimizer=o.kind,
            learning_rate=float(o.learning_rate), epsilon=float(o.epsilon), beta_1=float(o.beta_1), beta_2=float(o.beta_2),
            amsgrad=bool(o.amsgrad), initial_accumulator_value=float(o.initial_accumulator_value),
            momentum=float(o.momentum), nesterov=bool(o.nesterov), lr_scheduler=bool(o.lr_scheduler),
            decay_steps=float(o.decay_steps), decay_rate=float(o.decay_rate), step_offset=int(step_offset),
            use_btv=bool(self.use_BTV))
        for k, v in overrides.items():
            setattr(p, k, v)
        return p

    def _dropout_keep(self, n_copies: int):
        """Copy dropout (reference :47-53, :118): int(num_aug*copy_dropout) copies are masked out by a
        NumPy-shuffled boolean mask that the reference draws once, when loss_function is traced."""
        n_drop = int(self.num_aug * self.copy_dropout)
        if n_drop == 0:
            return None
        if self._keep_mask is None or len(self._keep_mask) != self.num_aug:
            mask = np.full(self.num_aug, fill_value=True)
            mask[:n_drop] = False
            np.random.shuffle(mask)
            self._keep_mask = mask
        if n_copies != self.num_aug:
            raise ValueError("copy dropout needs len(copies) == num_aug (tf.boolean_mask would fail in the reference)")
        return self._keep_mask.astype(np.uint8)

    def _check_optimizer(self):
        if self.optimizer is None:
            raise Exception(
                "You must provide an instance of the Optimizer class to compute the augmented SR")

    def _check_sizes(self, h, w):
        H, W = self.output_size
        if H % h or W % w or H // h != W // w or (H // h) % 2:
            raise NotImplementedError(f"libasr implements output_size == feature map size times an even integer; got {(h, w)} -> {(H, W)}")

    # ---- reference API ---------------------------------------------------------------------------------------
    def augmented_superresolution(self, augmented_copies, angles, shifts):
        """reference :102-137 -> (target_image [H,W,1] float32 ndarray, loss float)."""
        self._check_optimizer()
        stack = _as_device_stack(augmented_copies)
        n, h, w = stack.shape
        self._check_sizes(h, w)
        keep = self._dropout_keep(n)
        params = self._solve_params(self.optimizer.iterations)
        res = _lib.solve_batched(stack[None], np.asarray(angles, np.float32)[None], np.asarray(shifts, np.float32)[None],
                                 params, keep=None if keep is None else keep[None], want_loss=True, output_size=self.output_size,
                                 loss_every=10 if self.verbose else 0)
        x, loss = res[0], res[1]
        self.optimizer.iterations += int(self.num_iter)     # Keras' shared step counter keeps counting
        out = x[0].cpu().numpy()[..., None]
        loss = float(loss[0].item())
        if self.verbose:   # reference :129-130: every tenth iteration and the last one (printed after the solve here, not during it)
            trace = res[2][0].cpu().numpy()
            for i in range(int(self.num_iter)):
                if i % 10 == 0 or i == self.num_iter - 1:
                    v = loss if i == self.num_iter - 1 else float(trace[i // 10])
                    print(f"{i + 1}/{self.num_iter} -- loss = {v}")
        return out, loss

    def max_superresolution(self, augmented_copies, angles, shifts):
        """reference :139-149 -> ([H,W,1] ndarray, None)."""
        return self._backproject(augmented_copies, angles, shifts, "max"), None

    def mean_superresolution(self, augmented_copies, angles, shifts):
        """reference :151-161 -> ([H,W,1] ndarray, None)."""
        return self._backproject(augmented_copies, angles, shifts, "mean"), None

    def _backproject(self, augmented_copies, angles, shifts, mode):
        stack = _as_device_stack(augmented_copies)
        out = self.backproject_batched(stack[None], np.asarray(angles, np.float32)[None], np.asarray(shifts, np.float32)[None], mode)
        return out[0].cpu().numpy()[..., None]

    # ---- batched extensions (device in, device out) ---------------------------------------------------------------
    def augmented_superresolution_batched(self, copies, angles, shifts, params_list=None, keep=None, want_loss=False,
                                          advance_iterations=True):
        """Solve B images in one call.  copies: CUDA [B,N,h,w]; angles [B,N]; shifts [B,N,2].
        Image j starts at optimizer step `iterations + j*num_iter`, as the reference's sequential loop over
        images with one shared optimizer would (SR_single_class.py:66-107).  Pass `params_list` (one
        SolveParams per image) to solve a hyper-parameter grid instead (sweep_script.py:88-130)."""
        self._check_optimizer()
        B, n, h, w = copies.shape
        self._check_sizes(h, w)
        if params_list is None:
            base = self.optimizer.iterations
            params_list = [self._solve_params(base + j * int(self.num_iter)) for j in range(B)]
            if advance_iterations:
                self.optimizer.iterations += B * int(self.num_iter)
        return _lib.solve_batched(copies, angles, shifts, params_list, keep=keep, want_loss=want_loss, output_size=self.output_size)

    def backproject_batched(self, copies, angles, shifts, mode):
        torch = _lib._torch()
        L = _lib.lib()
        B, n, h, w = copies.shape
        H, W = self.output_size
        out = torch.empty((B, H, W), dtype=torch.float32, device=copies.device)
        ang = _lib._host_f32(angles, (B, n))
        shf = _lib._host_f32(shifts, (B, n, 2))
        need = C.c_size_t()
        _lib.check(L.asr_backproject_workspace_bytes(B, n, C.byref(need)))
        ws = _lib._aux_ws.get(need.value, copies.device)
        with torch.cuda.device(copies.device):
            _lib.check(L.asr_backproject_batched_ws(0 if mode == "max" else 1, copies.data_ptr(), ang.ctypes.data_as(_lib._fp),
                                                    shf.ctypes.data_as(_lib._fp), B, n, h, w, H, W, out.data_ptr(),
                                                    ws.data_ptr(), ws.numel(), _lib._stream_ptr(torch)))
        return out
