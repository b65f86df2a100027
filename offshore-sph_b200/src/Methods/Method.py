"""Method base class (reference src/Methods/Method.py:5-36)."""
import abc


class Method(abc.ABC):
    @abc.abstractmethod
    def initialize(self, pA):
        raise Exception('No implemented!')

    @abc.abstractmethod
    def compute_speed_of_sound(self, pA):
        raise Exception('No implemented!')

    @abc.abstractmethod
    def compute_pressure(self, pA):
        raise Exception('No implemented!')
