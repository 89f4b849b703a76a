"""B200-native CNN2 modulation-recognition hot path."""
__version__ = "0.1.0"
