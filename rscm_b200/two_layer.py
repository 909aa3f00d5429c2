"""Mirror of ``rscm._lib.two_layer`` (python/rscm/_lib/two_layer.pyi)."""

from . import _ffi
from ._builders import ComponentBuilder

__all__ = ["TwoLayerBuilder"]


class TwoLayerBuilder(ComponentBuilder):
    """TwoLayerParameters — crates/rscm-two-layer/src/component.rs:38-90 (all fields required)."""

    KIND = _ffi.TWO_LAYER
    TYPE_NAME = "TwoLayer"
    FIELDS = (
        ("lambda0", None),
        ("a", None),
        ("efficacy", None),
        ("eta", None),
        ("heat_capacity_surface", None),
        ("heat_capacity_deep", None),
    )
