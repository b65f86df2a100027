"""Post-processing stub with the constructor of reference src/Post/Plot.py.  Visualisation (pyqtgraph frames
-> ffmpeg) is outside the hot path this package replaces; the class exists so example scripts import."""


class Plot:
    def __init__(self, file: str, title: str = '', xmin=None, xmax=None, ymin=None, ymax=None, **kw):
        self.file, self.title = file, title
        self.limits = (xmin, xmax, ymin, ymax)

    def save(self, location: str):
        raise NotImplementedError('plotting is not part of the B200 WCSPH package; read %s with your own tools'
                                  % self.file)
