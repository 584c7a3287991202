"""Engine part of __graft_entry__.smoke(): grows with the engine (UNet / sampler checks vs the oracle)."""


def run(dev) -> None:
    return None
