"""placeholder: the reference imports jax_md.energy but the hot path does not call it"""
