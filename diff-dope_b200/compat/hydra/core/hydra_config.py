class HydraConfig:
    """Singleton holding the run's hydra config node (runtime.output_dir is what the reference reads)."""

    _cfg = None

    @classmethod
    def _set(cls, cfg):
        cls._cfg = cfg

    @classmethod
    def get(cls):
        if cls._cfg is None:
            raise ValueError("HydraConfig was not set")
        return cls._cfg

    @classmethod
    def initialized(cls):
        return cls._cfg is not None
