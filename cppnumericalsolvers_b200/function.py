"""Objective descriptors: the host-side mirror of the reference's function model
(include/cppoptlib/function_base.h) with a batch axis.

A Python object here only *names* a device functor that was compiled into
libcno.so (csrc/cno_functors.cuh); evaluation never leaves the GPU.
`BatchedFunctionState` is `FunctionState` (function_base.h:298-332) with a
leading batch dimension: x [B, d], value [B], gradient [B, d].
"""
from __future__ import annotations

import enum
from dataclasses import dataclass
from typing import Optional

import torch

from . import _lib


class DifferentiabilityMode(enum.IntEnum):  # function_base.h:42-46
    NONE = 0
    First = 1
    Second = 2


_DTYPES = {torch.float64: _lib.F64, torch.float32: _lib.F32}


def default_policy(dtype: torch.dtype) -> int:
    """fp64 kernels reduce on the FP64 tensor core (DMMA tree), fp32 kernels with
    the shuffle butterfly -- DESIGN.md "Arithmetic specification"."""
    return _lib.POLICY_DMMA_TREE if dtype == torch.float64 else _lib.POLICY_WARP_TREE


@dataclass
class Function:
    """Base descriptor (FunctionCRTP's static members, function_base.h:96-102)."""
    Dimension: int
    ScalarType: torch.dtype = torch.float64
    Differentiability: DifferentiabilityMode = DifferentiabilityMode.First
    family: int = -1
    n: int = 0
    param: float = 0.0
    data: Optional[torch.Tensor] = None  # per-instance data [B, stride]
    policy: Optional[int] = None  # None = the policy the kernels of this dtype implement

    def problem(self) -> _lib.Problem:
        data_ptr, stride = None, 0
        if self.data is not None:
            assert self.data.dtype == self.ScalarType and self.data.is_contiguous()
            data_ptr, stride = self.data.data_ptr(), self.data.shape[1]
        policy = self.policy
        if policy is None:
            policy = default_policy(self.ScalarType)
        mode = 2 if self.Differentiability == DifferentiabilityMode.Second else 0
        return _lib.Problem(self.family, _DTYPES[self.ScalarType], self.Dimension, self.n,
                            float(self.param), data_ptr, stride, policy, mode)


def Rosenbrock(d: int, dtype: torch.dtype = torch.float64) -> Function:
    """Chained Rosenbrock; d = 2 is src/test/verify.cc:58-69."""
    return Function(d, dtype, DifferentiabilityMode.First, _lib.FN_ROSENBROCK)


def RosenbrockFull(d: int, dtype: torch.dtype = torch.float64) -> Function:
    """Second-mode chained Rosenbrock (src/test/verify.cc:81-99): NewtonDescent, or Lbfgs
    with its diagonal preconditioner (solver/lbfgs.h:116-139)."""
    return Function(d, dtype, DifferentiabilityMode.Second, _lib.FN_ROSENBROCK)


def DiagQuadratic(dtype: torch.dtype = torch.float64) -> Function:
    """5 x0^2 + 100 x1^2 + 5 (Dockerfile.test:21-29)."""
    return Function(2, dtype, DifferentiabilityMode.First, _lib.FN_DIAG_QUADRATIC)


def HalfSquaredNorm(d: int, dtype: torch.dtype = torch.float64) -> Function:
    """0.5 ||x||^2 (src/test/augmented_lagrangian_test.cc:123-130)."""
    return Function(d, dtype, DifferentiabilityMode.First, _lib.FN_HALF_SQUARED_NORM)


def Logistic(data: torch.Tensor, n: int, d: int, lam: float) -> Function:
    """Batched logistic regression; data[b] = [X (n x d row-major) | y (n)]."""
    return Function(d, data.dtype, DifferentiabilityMode.First, _lib.FN_LOGISTIC, n=n,
                    param=lam, data=data)


def DenseQuadratic(data: torch.Tensor, d: int, policy: Optional[int] = None) -> Function:
    """0.5 x'Ax - b'x; data[b] = [A (d x d col-major, bitwise symmetric) | b (d)]; Second mode.
    policy = POLICY_DMMA_LU (d = 64 fp64): NewtonDescent factors the Hessian with fused multiply-subtracts, the
    trailing update of the blocked elimination on the FP64 tensor core (include/cno.h, csrc/cno_newton_dmma.cuh)."""
    return Function(d, data.dtype, DifferentiabilityMode.Second, _lib.FN_DENSE_QUADRATIC,
                    data=data, policy=policy)


def DenseQuadraticFirst(data: torch.Tensor, d: int) -> Function:
    """The same objective as a First-mode function (value and gradient only): for Lbfgs and for
    AugmentedLagrangian's batched "quadratic objective, affine / ball constraints" problems."""
    return Function(d, data.dtype, DifferentiabilityMode.First, _lib.FN_DENSE_QUADRATIC, data=data)


@dataclass
class BatchedFunctionState:
    """FunctionState with a batch axis (function_base.h:298-332)."""
    x: torch.Tensor
    value: Optional[torch.Tensor] = None
    gradient: Optional[torch.Tensor] = None

    @property
    def batch(self) -> int:
        return self.x.shape[0]
