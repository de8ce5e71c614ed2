"""medicalseg.utils.logger (reference logger.py:24-50): `logger.info / warning / error / debug(message)` print one
timestamped, levelled line on local rank 0 - the format core.train's own log lines use."""
import os
import sys
import time

_NAMES = ("ERROR", "WARNING", "INFO", "DEBUG")
log_level = 2  # messages above this verbosity are dropped (2 = INFO, as the reference)


def _is_main_process():
    return os.environ.get("LOCAL_RANK", "0") in ("0", "")


def log(level=2, message=""):
    if level > log_level or not _is_main_process():
        return
    sys.stdout.write("%s [%s]\t%s\n" % (time.strftime("%Y-%m-%d %H:%M:%S"), _NAMES[level], message))
    sys.stdout.flush()


def _at(level):
    def emit(message=""):
        log(level, message)
    return emit


error, warning, info, debug = _at(0), _at(1), _at(2), _at(3)
