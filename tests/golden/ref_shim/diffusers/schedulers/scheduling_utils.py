from enum import Enum


class KarrasDiffusionSchedulers(Enum):
    EulerDiscreteScheduler = 3


class SchedulerMixin:
    config_name = "scheduler_config.json"
    _compatibles = []
    has_compatibles = True
