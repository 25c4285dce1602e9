"""placeholder: the reference imports jax_md.simulate but the hot path does not call it"""
