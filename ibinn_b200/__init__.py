"""ibinn_b200 -- B200-native engine for the IB-INN hot path."""
__version__ = '0.1.0'
