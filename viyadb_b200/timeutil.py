"""Host-side time arithmetic: mirror of util::Duration / time literal parsing of the reference.

Only plan *constants* are computed here (rollup boundaries, filter literals); per-row truncation
happens on the device (csrc/kernels.cuh: rollup_value).

Reference: src/util/time.h:27-137, src/util/time.cc:49-83 (Duration::add_to via gmtime_r/timegm),
src/codegen/query/filter.cc:167-196 (time literal decoding).
"""
import calendar
import re
import time as _time

TIME_UNITS = ["year", "month", "week", "day", "hour", "minute", "second"]  # util::TimeUnit order


def time_unit_by_name(name):
    """src/util/time.cc:52-59."""
    if name in TIME_UNITS:
        return TIME_UNITS.index(name)
    raise ValueError("Unsupported time unit: " + name)


def _timegm_normalised(year, mon0, mday, hour, minute, sec):
    """timegm() on a struct tm whose fields may be out of range (glibc normalises them)."""
    year += mon0 // 12
    mon0 %= 12
    days = calendar.timegm((year, mon0 + 1, 1, 0, 0, 0)) // 86400 if year >= 1970 else _days_from_civil(year, mon0 + 1, 1)
    return (days + (mday - 1)) * 86400 + hour * 3600 + minute * 60 + sec


def _days_from_civil(y, m, d):
    y -= m <= 2
    era = (y if y >= 0 else y - 399) // 400
    yoe = y - era * 400
    doy = (153 * (m + (-3 if m > 2 else 9)) + 2) // 5 + d - 1
    doe = yoe * 365 + yoe // 4 - yoe // 100 + doy
    return era * 146097 + doe - 719468


class Duration:
    """util::Duration (src/util/time.h:35-56, time.cc:61-105): '<n> <unit>s'."""

    def __init__(self, desc):
        parts = desc.split()
        try:
            n = int(parts[0])
            unit_name = parts[1]
        except (IndexError, ValueError):
            raise ValueError("Wrong duration description: " + desc)
        if n <= 0:
            raise ValueError("Wrong duration description: " + desc)
        self.count = n
        self.time_unit = time_unit_by_name(unit_name[:-1])  # tu_name.pop_back()

    def add_to(self, timestamp, sign):
        """uint32 overload: calendar field arithmetic in UTC, then timegm; wraps to uint32."""
        tm = _time.gmtime(int(timestamp) & 0xFFFFFFFF)
        y, mo, d, h, mi, s = tm.tm_year, tm.tm_mon - 1, tm.tm_mday, tm.tm_hour, tm.tm_min, tm.tm_sec
        delta = sign * self.count
        u = self.time_unit
        if u == 0:
            y += delta
        elif u == 1:
            mo += delta
        elif u == 2:
            d += 7 * delta
        elif u == 3:
            d += delta
        elif u == 4:
            h += delta
        elif u == 5:
            mi += delta
        elif u == 6:
            s += delta
        else:
            raise RuntimeError("Unsupported duration")
        return _timegm_normalised(y, mo, d, h, mi, s) & 0xFFFFFFFF

    def add_to_micro(self, timestamp, sign):
        """uint64 overload (time.cc:103-105)."""
        return (self.add_to((timestamp // 1000000) & 0xFFFFFFFF, sign) * 1000000) & 0xFFFFFFFFFFFFFFFF

    def seconds_from_epoch(self):
        """Ordering key of Duration::operator> (time.h:48-50)."""
        return self.add_to(0, 1)


_DATE_TIME = re.compile(r"^\s*(\d{1,4})-(\d{1,2})-(\d{1,2})\s+(\d{1,2}):(\d{1,2}):(\d{1,2})(.*)$")
_DATE = re.compile(r"^\s*(\d{1,4})-(\d{1,2})-(\d{1,2})$")


def parse_time_literal(value, micro):
    """ValueDecoder::Visit(TimeDimension) (src/codegen/query/filter.cc:167-196).

    All-digit strings are raw timestamps; otherwise '%Y-%m-%d %T' (optionally '.<micros>' for
    microtime) or '%Y-%m-%d'. Returns the integer in the column's unit; raises ValueError with the
    reference's message when nothing matches (or the result is 0, which the reference also rejects).
    """
    if value != "" and all(c.isdigit() for c in value):
        return int(value) & (0xFFFFFFFFFFFFFFFF if micro else 0xFFFFFFFF)
    if value == "":
        # std::all_of on an empty range is true -> Parse("") -> std::stoul throws invalid_argument
        raise ValueError("stoul")
    mult = 1000000 if micro else 1
    ts = 0
    m = _DATE_TIME.match(value)
    if m:
        y, mo, d, h, mi, s, rest = m.groups()
        base = _timegm_normalised(int(y), int(mo) - 1, int(d), int(h), int(mi), int(s))
        if rest == "":
            ts = base * mult
        elif micro and rest.startswith("."):
            # the reference calls std::stoul(r) with r pointing AT the dot (filter.cc:178), which
            # throws std::invalid_argument("stoul") for every input: same error here
            raise ValueError("stoul")
    else:
        m = _DATE.match(value)
        if m:
            y, mo, d = m.groups()
            ts = _timegm_normalised(int(y), int(mo) - 1, int(d), 0, 0, 0) * mult
    if ts > 0:
        return ts & (0xFFFFFFFFFFFFFFFF if micro else 0xFFFFFFFF)
    raise ValueError("Unrecognized time format: " + value)
