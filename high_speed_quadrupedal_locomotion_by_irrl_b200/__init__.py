"""B200-native batched bp5 quadruped environment + LSTM act path (drop-in for IRRL's raisimGym hot path)."""
__version__ = "0.1.0"
