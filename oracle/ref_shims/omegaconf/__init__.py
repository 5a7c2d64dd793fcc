"""Import shim (test infrastructure only): the reference's lib/model/training.py and lib/utils/runner_utils.py import
``omegaconf`` at module top level for type annotations and config conversion; the evaluation episode
(training.py:54-96) uses none of it."""


class DictConfig(dict):
    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError as e:
            raise AttributeError(k) from e


class ListConfig(list):
    pass


class OmegaConf:
    @staticmethod
    def to_container(cfg, resolve=True, **kwargs):
        return dict(cfg) if isinstance(cfg, dict) else cfg

    @staticmethod
    def create(obj=None, **kwargs):
        return DictConfig(obj or {})

    @staticmethod
    def to_yaml(cfg, **kwargs):
        import yaml
        return yaml.safe_dump(dict(cfg))

    @staticmethod
    def save(config, f, **kwargs):
        open(f, "w").write(OmegaConf.to_yaml(config))


def open_dict(cfg):
    import contextlib
    return contextlib.nullcontext(cfg)
