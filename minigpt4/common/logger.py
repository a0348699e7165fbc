"""Training-loop meters with the reference's interface (minigpt4/common/logger.py: SmoothedValue :19-80, MetricLogger
:82-181, AttrDict :184, setup_logger :190). Own implementation: a meter keeps a bounded window plus running totals; the
cross-rank synchronisation reduces (count, total) over the process group, as the reference does."""
import collections
import datetime
import logging
import time

import torch
import torch.distributed as dist

from minigpt4.common import dist_utils


class SmoothedValue:
    def __init__(self, window_size=20, fmt=None):
        self.window = collections.deque(maxlen=window_size)
        self.total, self.count = 0.0, 0
        self.fmt = fmt or "{median:.4f} ({global_avg:.4f})"

    def update(self, value, n=1):
        self.window.append(float(value))
        self.count += n
        self.total += float(value) * n

    def synchronize_between_processes(self):
        """Only count / total are reduced; the window stays local (reference :37-48)."""
        if not dist_utils.is_dist_avail_and_initialized():
            return
        dev = "cuda" if (torch.cuda.is_available() and dist.get_backend() == "nccl") else "cpu"
        t = torch.tensor([self.count, self.total], dtype=torch.float64, device=dev)
        dist.barrier()
        dist.all_reduce(t)
        self.count, self.total = int(t[0].item()), float(t[1].item())

    @property
    def median(self):
        return float(torch.tensor(list(self.window)).median()) if self.window else 0.0

    @property
    def avg(self):
        return sum(self.window) / len(self.window) if self.window else 0.0

    @property
    def global_avg(self):
        return self.total / max(self.count, 1)

    @property
    def max(self):
        return max(self.window) if self.window else 0.0

    @property
    def value(self):
        return self.window[-1] if self.window else 0.0

    def __str__(self):
        return self.fmt.format(median=self.median, avg=self.avg, global_avg=self.global_avg, max=self.max, value=self.value)


class MetricLogger:
    def __init__(self, delimiter="\t"):
        self.meters = collections.defaultdict(SmoothedValue)
        self.delimiter = delimiter

    def update(self, **kwargs):
        for name, v in kwargs.items():
            if isinstance(v, torch.Tensor):
                v = v.item()
            assert isinstance(v, (float, int)), "meter %s got %r" % (name, type(v))
            self.meters[name].update(v)

    def __getattr__(self, attr):
        meters = self.__dict__.get("meters", {})
        if attr in meters:
            return meters[attr]
        raise AttributeError("'%s' object has no attribute '%s'" % (type(self).__name__, attr))

    def __str__(self):
        return self.delimiter.join("%s: %s" % (n, m) for n, m in self.meters.items())

    def global_avg(self):
        return self.delimiter.join("%s: %.4f" % (n, m.global_avg) for n, m in self.meters.items())

    def synchronize_between_processes(self):
        for m in self.meters.values():
            m.synchronize_between_processes()

    def add_meter(self, name, meter):
        self.meters[name] = meter

    def log_every(self, iterable, print_freq, header=None):
        """Yield from `iterable`, printing the meters, iteration / data time and an ETA every `print_freq` items."""
        header = header or ""
        n = len(iterable)
        iter_t, data_t = SmoothedValue(fmt="{avg:.4f}"), SmoothedValue(fmt="{avg:.4f}")
        width = len(str(n))
        t_start = t_last = time.time()
        for i, obj in enumerate(iterable):
            data_t.update(time.time() - t_last)
            yield obj
            iter_t.update(time.time() - t_last)
            if i % print_freq == 0 or i == n - 1:
                eta = datetime.timedelta(seconds=int(iter_t.global_avg * (n - i)))
                msg = [header, "[%*d/%d]" % (width, i, n), "eta: %s" % eta, str(self), "time: %s" % iter_t, "data: %s" % data_t]
                if torch.cuda.is_available():
                    msg.append("max mem: %.0f" % (torch.cuda.max_memory_allocated() / (1024.0 * 1024.0)))
                print(self.delimiter.join(msg))
            t_last = time.time()
        total = time.time() - t_start
        print("%s Total time: %s (%.4f s / it)" % (header, datetime.timedelta(seconds=int(total)), total / max(n, 1)))


class AttrDict(dict):
    def __init__(self, *args, **kwargs):
        super().__init__(*args, **kwargs)
        self.__dict__ = self


def setup_logger():
    logging.basicConfig(level=logging.INFO if dist_utils.is_main_process() else logging.WARN,
                        format="%(asctime)s [%(levelname)s] %(message)s", handlers=[logging.StreamHandler()])
