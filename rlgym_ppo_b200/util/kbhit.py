"""Non-blocking keyboard poll for the p / c / q keys of Learner._learn (learner.py:247, 309-321).

The reference's KBHit calls termios.tcgetattr(stdin) unconditionally (kbhit.py:45-47) and therefore raises without a
TTY.  The learner must also run headless (gpurun, CI, torchrun ranks), so this one degrades to "no key pressed".
"""
import atexit
import os
import sys
from select import select


class KBHit(object):
    def __init__(self):
        self._fd = None
        self._old = None
        if os.name == "nt":
            return
        try:
            import termios
            if not sys.stdin.isatty():
                return
            self._fd = sys.stdin.fileno()
            self._old = termios.tcgetattr(self._fd)
            new = termios.tcgetattr(self._fd)
            new[3] = new[3] & ~termios.ICANON & ~termios.ECHO
            termios.tcsetattr(self._fd, termios.TCSAFLUSH, new)
            atexit.register(self.set_normal_term)
        except Exception:
            self._fd = None

    def set_normal_term(self):
        if self._fd is not None and self._old is not None:
            import termios
            try:
                termios.tcsetattr(self._fd, termios.TCSAFLUSH, self._old)
            except Exception:
                pass

    def getch(self):
        if os.name == "nt":
            import msvcrt
            return msvcrt.getch().decode("utf-8")
        return sys.stdin.read(1) if self._fd is not None else ""

    def kbhit(self):
        if os.name == "nt":
            import msvcrt
            return msvcrt.kbhit()
        if self._fd is None:
            return False
        ready, _, _ = select([sys.stdin], [], [], 0)
        return ready != []
