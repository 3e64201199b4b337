"""B200-native implementation of empanada's panoptic inference hot path (drop-in for the
`empanada_napari.inference` engines). See DESIGN.md."""
__version__ = "0.1.0"
