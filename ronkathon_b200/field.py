"""Host-side mirror of ronkathon's `PrimeField<const P: usize>` and the `Field` / `FiniteField`
traits (src/algebra/field/mod.rs:17-76, src/algebra/field/prime/{mod,arithmetic}.rs).

Every arithmetic operation — including single-element ones — is executed by the CUDA kernels in
libronk_b200.so through the C ABI; this module contains no field arithmetic of its own.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from ._lib import GOLDILOCKS, RonkPanic


def _one(x: int) -> np.ndarray:
    return np.array([x], dtype=np.uint64)


class _FieldMeta(type):
    def __repr__(cls):
        return f"PrimeField<{cls.ORDER}>"

    @property
    def ZERO(cls):
        return cls(0)

    @property
    def ONE(cls):
        return cls(1 % cls.ORDER)

    @property
    def PRIMITIVE_ELEMENT(cls):
        """FiniteField::PRIMITIVE_ELEMENT (prime/mod.rs:87-90)."""
        g = C.c_uint64()
        rc = _lib.lib().ronk_field_generator(cls.ORDER, C.byref(g))
        if rc != 0:
            raise RonkPanic(rc, "generator not found")
        return cls(g.value)


class _Element(metaclass=_FieldMeta):
    ORDER = 0
    __slots__ = ("value",)

    def __init__(self, value: int):
        # PrimeField::new (prime/mod.rs:48-51): value % P.  (Reduction of a Python int on
        # construction is representation, not field arithmetic.)
        self.value = int(value) % self.ORDER

    # --- constructors / trait items ---------------------------------------------------------
    @classmethod
    def new(cls, value: int):
        return cls(value)

    @classmethod
    def primitive_root_of_unity(cls, n: int):
        """FiniteField::primitive_root_of_unity (field/mod.rs:70-75); panics if n ∤ P-1."""
        out = C.c_uint64()
        rc = _lib.lib().ronk_root_of_unity(cls.ORDER, cls.PRIMITIVE_ELEMENT.value, n, C.byref(out))
        if rc != 0:
            raise RonkPanic(rc, "n must divide p^q - 1")
        return cls(out.value)

    # --- arithmetic through the kernels --------------------------------------------------------
    def _bin(self, op: int, rhs):
        rhs = self._coerce(rhs)
        out = np.empty(1, dtype=np.uint64)
        _lib.default_context().call("ronk_field_binop_u64_host", op, self.ORDER, _lib._ptr(_one(self.value)),
                                    _lib._ptr(_one(rhs.value)), _lib._ptr(out), 1)
        return type(self)(int(out[0]))

    def _coerce(self, x):
        if isinstance(x, _Element):
            if x.ORDER != self.ORDER:
                raise TypeError("mismatched fields")
            return x
        return type(self)(x)

    def __add__(self, rhs): return self._bin(0, rhs)
    def __sub__(self, rhs): return self._bin(1, rhs)
    def __mul__(self, rhs): return self._bin(2, rhs)
    def __truediv__(self, rhs): return self._bin(3, rhs)  # prime/arithmetic.rs:54 (panics on 0)
    def div(self, rhs): return self._bin(3, rhs)

    def __neg__(self):
        out = np.empty(1, dtype=np.uint64)
        _lib.default_context().call("ronk_field_unop_u64_host", 0, self.ORDER, _lib._ptr(_one(self.value)),
                                    _lib._ptr(out), 1)
        return type(self)(int(out[0]))

    def inverse(self):
        """Field::inverse (prime/mod.rs:62-72): None for zero."""
        out = np.empty(1, dtype=np.uint64)
        try:
            _lib.default_context().call("ronk_field_unop_u64_host", 1, self.ORDER, _lib._ptr(_one(self.value)),
                                        _lib._ptr(out), 1)
        except RonkPanic:
            return None
        return type(self)(int(out[0]))

    def pow(self, power: int):
        """Field::pow (prime/mod.rs:74-84)."""
        out = np.empty(1, dtype=np.uint64)
        _lib.default_context().call("ronk_field_pow_u64_host", self.ORDER, _lib._ptr(_one(self.value)), int(power),
                                    _lib._ptr(out), 1)
        return type(self)(int(out[0]))

    def __mod__(self, rhs):  # Rem (prime/arithmetic.rs:70): self - (self / rhs) * rhs
        rhs = self._coerce(rhs)
        return self - (self / rhs) * rhs

    def __eq__(self, other):
        if isinstance(other, _Element):
            return self.ORDER == other.ORDER and self.value == other.value
        return NotImplemented

    def __hash__(self):
        return hash((self.ORDER, self.value))

    def __int__(self):
        return self.value

    def __repr__(self):
        return f"{self.value}"


_cache: dict[int, type] = {}


def PrimeField(p: int) -> type:
    """`PrimeField::<P>` — returns the element class for modulus `p`."""
    if p not in _cache:
        _cache[p] = _FieldMeta(f"PrimeField_{p}", (_Element,), {"ORDER": int(p), "__slots__": ()})
    return _cache[p]


PlutoBaseField = PrimeField(101)    # prime/mod.rs:27
PlutoScalarField = PrimeField(17)   # prime/mod.rs:31
GoldilocksField = PrimeField(GOLDILOCKS)  # the 64-bit instantiation


def sum_field(items, field):
    """Sum (prime/arithmetic.rs:13-17): reduce(+) or ZERO."""
    acc = None
    for x in items:
        acc = x if acc is None else acc + x
    return field.ZERO if acc is None else acc
