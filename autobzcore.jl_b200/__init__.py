"""autobz_b200 — B200-native hot path of AutoBZCore.jl behind the reference's solve/init API."""
from . import _lib, synthetic  # noqa: F401
