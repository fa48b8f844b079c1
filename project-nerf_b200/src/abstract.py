"""Plugin contracts of the hot path (mirrors reference src/abstract.py:4-26).

Encoders map coordinates to features and advertise their width through
``out_dim``; decoders map features to radiance/density (or a displacement).
"""
from abc import ABC, abstractmethod

from torch import nn


class BaseRepresentation(nn.Module, ABC):
    """coords [P, D] -> features [P, out_dim]."""

    @abstractmethod
    def forward(self, x):
        raise NotImplementedError

    @property
    @abstractmethod
    def out_dim(self) -> int:
        """feature width, so decoders can size their first layer."""
        raise NotImplementedError


class BaseDecoder(nn.Module, ABC):
    """features -> outputs (rgb, sigma, displacement ...)."""

    @abstractmethod
    def forward(self, x):
        raise NotImplementedError
