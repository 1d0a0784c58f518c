"""ignis_b200: B200-native device for the Ignis path-tracing hot path."""
__version__ = "0.1.0"
