"""`--flag / --no-flag` boolean action and the `parser.add` alias the plugins' `get_args` hooks expect
(the reference's utils/argparse_utils.py contract)."""
import argparse


class StoreBool(argparse.Action):
    """Registers both `--name` (True) and `--no-name` (False) for one destination."""

    def __init__(self, option_strings, dest, default=None, required=False, help="", **_ignored):
        if len(option_strings) != 1 or not option_strings[0].startswith('--'):
            raise ValueError("store_bool expects exactly one long option")
        name = option_strings[0][2:]
        super().__init__([f'--{name}', f'--no-{name}'], dest=dest, nargs=0, default=default, required=required,
                         help=f'{help}Use "--{name}" for True, "--no-{name}" for False')

    def __call__(self, parser, namespace, values, option_string=None):
        setattr(namespace, self.dest, not option_string.startswith('--no-'))


class MyArgumentParser(argparse.ArgumentParser):
    def __init__(self, **kwargs):
        super().__init__(**kwargs)
        self.register('action', 'store_bool', StoreBool)

    def add(self, *args, **kwargs):
        return self.add_argument(*args, **kwargs)
