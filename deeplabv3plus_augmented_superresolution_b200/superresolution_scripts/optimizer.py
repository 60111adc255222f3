"""Drop-in for superresolution_scripts/optimizer.py (reference :4-52) without TensorFlow.

The reference wraps a Keras optimizer; here the object only records the hyper-parameters and the one
piece of state that outlives a solve in the reference: Keras' `optimizer.iterations`, which keeps
counting across images because one Optimizer is shared by the whole run (SR_single_class.py:66-70,
SURVEY.md Appendix B-1).  The arithmetic of every optimizer lives in k_gradient_update
(csrc/asr_solve.cu); slots (m, v, vhat, accumulators) are re-created per solve exactly as a fresh
tf.Variable gets fresh slots.
"""
from __future__ import annotations

import numpy as np

_KNOWN = ("adadelta", "adagrad", "adamax", "sgd")


class Optimizer:
    def __init__(self, optimizer="adam", learning_rate=1e-3,
                 epsilon=1e-7, beta_1=.9, beta_2=.999, amsgrad=False,
                 initial_accumulator_value=.1, momentum=.0, nesterov=False,
                 lr_scheduler=False, decay_steps=.5, decay_rate=100) -> None:
        self.learning_rate = learning_rate
        self.epsilon = epsilon
        self.beta_1 = beta_1
        self.beta_2 = beta_2
        self.amsgrad = amsgrad
        self.initial_accumulator_value = initial_accumulator_value
        self.momentum = momentum
        self.nesterov = nesterov
        self.decay_steps = decay_steps
        self.decay_rate = decay_rate
        # the reference falls through to Adam for any unknown name (optimizer.py:36-41)
        self.kind = optimizer if optimizer in _KNOWN else "adam"
        self.lr_scheduler = bool(lr_scheduler)
        # Keras `optimizer.iterations`: number of apply_gradients calls so far, never reset by the reference
        self.iterations = 0
        self.current_learning_rate = np.float32(learning_rate)

    def lr_decay(self, iteration):
        """ExponentialDecay(lr0, decay_steps, decay_rate)(iteration) in fp32 (reference :50-52).
        Kept for API compatibility; the solve evaluates the same schedule inside libasr."""
        p = np.float32(iteration) / np.float32(self.decay_steps)
        self.current_learning_rate = np.float32(self.learning_rate) * np.power(np.float32(self.decay_rate), p,
                                                                              dtype=np.float32)
        return self.current_learning_rate
