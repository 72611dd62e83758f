"""Process-wide 'rl' logger with the reference's format and daily-rotating file (rl/utils/logging.py:9-48)."""
import logging
from logging.handlers import TimedRotatingFileHandler

logger = None
_FORMAT = '%(asctime).19s [%(levelname)s] %(message)s'


def init_logging(save_path=None, add_std_err=True):
    global logger
    if logger is not None:
        return logger
    lg = logging.getLogger('rl')
    lg.setLevel(logging.DEBUG)
    handlers = []
    if save_path is not None:
        handlers.append(TimedRotatingFileHandler(save_path + '.log', when='D'))
    if add_std_err:
        handlers.append(logging.StreamHandler())
    for hd in handlers:
        hd.setFormatter(logging.Formatter(_FORMAT))
        lg.addHandler(hd)
    logger = lg
    return logger


def log(*mesg):
    text = ' '.join(str(m) for m in mesg)
    if logger is None:
        print(text)
    else:
        logger.info(text)
