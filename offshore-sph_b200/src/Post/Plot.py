"""Post-processing stand-in with the constructor of reference src/Post/Plot.py.  Visualisation (pyqtgraph frames
-> ffmpeg) is outside the hot path this package replaces; the class exists so that unedited example scripts import
and run to their end (examples/Containment.py ships with plot = True): `save` reports where the data is and returns."""
import sys


class Plot:
    def __init__(self, file: str, title: str = '', xmin=None, xmax=None, ymin=None, ymax=None, **kw):
        self.file, self.title = file, title
        self.limits = (xmin, xmax, ymin, ymax)

    def save(self, location: str):
        print('Plot.save(%s): animation rendering is not part of the B200 WCSPH package; the solver output is in %s'
              % (location, self.file), file=sys.stderr)
        return None
