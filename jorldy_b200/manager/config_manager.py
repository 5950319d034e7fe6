"""Config loading with the reference's surface (jorldy/manager/config_manager.py:4-98): a config is a
python module with four dicts `agent / env / optim / train`, addressed by dotted path
(`config.ppo.cartpole`), wrapped in attribute-dicts, and overridable from the command line with
`--domain.key value` / `--domain.key=value` pairs (int -> float -> bool -> None -> str casting;
`None` removes the key).

Resolution order for the dotted path: an importable module (so an existing JORLDY `config/` directory
on sys.path runs unchanged), else the built-in tables in jorldy_b200/config/.
"""
import importlib
import os


def type_cast(var):
    for cast in (int, float):
        try:
            return cast(var)
        except (TypeError, ValueError):
            pass
    if var in ("True", "False"):
        return var == "True"
    return None if var == "None" else var


class CustomDict(dict):
    __getattr__ = dict.get
    __setattr__ = dict.__setitem__
    __delattr__ = dict.__delitem__
    __getitem__ = __getattr__

    def __init__(self, init_dict=None):
        super().__init__()
        self.update(init_dict or {})

    def __getstate__(self):
        return self.__dict__

    def __setstate__(self, d):
        self.__dict__.update(d)


def _load_module(config_path):
    try:
        return importlib.import_module(config_path)
    except ImportError:
        from .. import config as builtin
        return builtin.load(config_path)


class ConfigManager:
    DOMAINS = ("env", "agent", "optim", "train")

    def __init__(self, config_path, unknown_args=()):
        module = _load_module(config_path)
        self.config = CustomDict()
        self.config.agent = CustomDict(module.agent)
        self.config.optim = CustomDict(module.optim)
        self.config.env = CustomDict(module.env)
        self.config.train = CustomDict(module.train)
        self.unknown_update(list(unknown_args))

    def unknown_update(self, unknown_args):
        removals = []
        i = 0
        while i < len(unknown_args):
            query = unknown_args[i]
            assert "--" in query, "use -- before the optional argument."
            if "=" in query:
                key, value = query.strip("-").split("=")
            else:
                key = query.strip("-")
                i += 1
                assert i < len(unknown_args) and "--" not in unknown_args[i], "check command again."
                value = unknown_args[i]
            assert "." in key and key.split(".")[0] in self.DOMAINS, \
                "optional argument should include env, agent or train. ex)env.name"
            domain, key = key.split(".")
            value = type_cast(value)
            if value is None:
                removals.append((domain, key))
            else:
                self.config[domain][key] = value
            i += 1
        for domain, key in removals:
            self.config[domain].pop(key, None)

    def dump(self, dump_path):
        with open(os.path.join(dump_path, "config.py"), "w", encoding="utf-8") as f:
            f.write(f"### {self.config.agent.name} {self.config.env.name} config ###\n")
            for domain in self.config.keys():
                f.write(f"\n{domain} = {{\n")
                for key, value in self.config[domain].items():
                    value = f"'{value}'" if type(value) == str else value
                    f.write(f"\t'{key}': {value},\n")
                f.write("}\n")
