"""Import shim: only the names are needed (configs are loaded with yaml.safe_load in the oracle)."""


class ListConfig(list):
    pass


class DictConfig(dict):
    pass


class OmegaConf:
    @staticmethod
    def load(path):
        import yaml
        with open(path) as f:
            return yaml.safe_load(f)
