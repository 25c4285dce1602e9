"""placeholder: imported by the reference, not called on the hot path"""
