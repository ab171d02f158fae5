"""Loggers for the host side.

Behaviour kept from the reference (sykepic/utils/logger.py): records look like
`<time> - <logger name> - <level> - <message>` and the level comes from the `LOGLEVEL` environment variable (INFO when
unset).  The reference's optional YAML logging configuration belongs to its sync daemon and is not part of this path.
"""

import logging
import os
import threading

_RECORD_LAYOUT = " - ".join("{%s}" % field for field in ("asctime", "name", "levelname", "message"))
_once = threading.Lock()
_ready = False


def _configure_root():
    """Installs one stream handler on the root logger, unless the application has already configured logging."""
    root = logging.getLogger()
    level = os.environ.get("LOGLEVEL", "INFO")
    if not root.handlers:
        handler = logging.StreamHandler()
        handler.setFormatter(logging.Formatter(_RECORD_LAYOUT, style="{"))
        root.addHandler(handler)
    try:
        root.setLevel(level)
    except ValueError:  # an unknown level name: keep INFO rather than fail at import time
        root.setLevel(logging.INFO)


def get_logger(name):
    """Logger `name`; the first call configures the root logger (thread-safe: the pipeline's threads log too)."""
    global _ready
    if not _ready:
        with _once:
            if not _ready:
                _configure_root()
                _ready = True
    return logging.getLogger(name)
