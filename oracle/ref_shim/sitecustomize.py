"""Compatibility shim so the unmodified reference imports on Python 3.12 / PyYAML 6.

Test infrastructure only.  Put this directory on PYTHONPATH *before*
importing `/root/reference/demo1`:
  * `inspect.getargspec` was removed in Python 3.11 (used at demo1/sim.py:210-211);
  * `yaml.load` needs a Loader since PyYAML 6 (called bare at demo1/config.py:180).
Nothing here changes the reference's arithmetic or control flow.
"""
import inspect

if not hasattr(inspect, "getargspec"):
    def _getargspec(func):
        spec = inspect.getfullargspec(func)
        return (spec.args, spec.varargs, spec.varkw, spec.defaults)
    inspect.getargspec = _getargspec

try:
    import yaml as _yaml
    _orig_load = _yaml.load

    def _load(stream, Loader=_yaml.SafeLoader):
        return _orig_load(stream, Loader=Loader)
    _yaml.load = _load
except ImportError:  # pragma: no cover
    pass
