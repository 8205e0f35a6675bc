"""TEST INFRASTRUCTURE (oracle): restatement of the reference's recording file-name logic, line by line,
with the reference's own regular expressions.  Only tests/ may import this.

  MainViewModel.setFilesourceUri   ui/MainViewModel.kt:2034-2080   metadata from a file name
  Long.asStringWithUnit            ui/composable/HelperComposables.kt:168-179
  Recording.calculateFileName      database/RecordingDao.kt:87-90
  FileIQSource.getPacket           source/FileIQSource.java:305-341   whole packets only, rewind on repeat
"""
import re

HACKRF, RTLSDR, AIRSPY, HYDRASDR = range(4)
FORMAT_NAMES = ["HACKRF", "RTLSDR", "AIRSPY", "HYDRASDR"]


def _matches(pattern, s):
    return re.fullmatch(pattern, s, flags=re.DOTALL) is not None  # String.matches = whole-string match


def parse_name(filename, file_format, frequency, sample_rate):
    """Returns (file_format, frequency, sample_rate) after setFilesourceUri's extraction."""
    try:
        if (_matches(".*hackrf.*", filename) or _matches(".*HackRF.*", filename) or _matches(".*HACKRF.*", filename)
                or _matches(".*hackrfone.*", filename)):
            file_format = HACKRF
        if (_matches(".*rtlsdr.*", filename) or _matches(".*rtl-sdr.*", filename) or _matches(".*RTLSDR.*", filename)
                or _matches(".*RTL-SDR.*", filename)):
            file_format = RTLSDR
        if (_matches(".*airspy.*", filename) or _matches(".*Airspy.*", filename) or _matches(".*AIRSPY.*", filename)
                or _matches(".*AirSpy.*", filename)):
            file_format = AIRSPY
        if (_matches(".*hydrasdr.*", filename) or _matches(".*HydraSDR.*", filename) or _matches(".*HYDRASDR.*", filename)
                or _matches(".*HydraSdr.*", filename)):
            file_format = HYDRASDR

        def grab(units, scale, current):
            pat = r".*(_|-|\s)([0-9]+)(%s).*" % units
            if _matches(pat, filename):
                digits = re.sub(pat, r"\2", filename, count=1, flags=re.DOTALL)
                if len(digits) > 18:
                    raise ValueError("NumberFormatException")
                return int(digits) * scale
            return current

        sample_rate = grab("sps|Sps|SPS", 1, sample_rate)
        sample_rate = grab("ksps|Ksps|KSps|KSPS", 1000, sample_rate)
        sample_rate = grab("msps|Msps|MSps|MSPS", 1000000, sample_rate)
        frequency = grab("hz|Hz|HZ", 1, frequency)
        frequency = grab("khz|Khz|KHz|KHZ", 1000, frequency)
        frequency = grab("mhz|Mhz|MHz|MHZ", 1000000, frequency)
    except ValueError:
        pass  # the app logs and keeps what it has so far
    return file_format, frequency, sample_rate


def as_string_with_unit(value, unit):
    units = ["", "k", "M", "G", "T"]
    index = 0
    while value % 1000 == 0 and value >= 1000 and index < len(units) - 1:
        value //= 1000
        index += 1
    return "{:,}".format(value).replace(",", " ") + " " + units[index] + unit


def calculate_file_name(timestamp, name, file_format, frequency, sample_rate):
    return "%s_%s_%s_%s_%s.iq" % (timestamp, name, FORMAT_NAMES[file_format],
                                  as_string_with_unit(frequency, "Hz").replace(" ", ""),
                                  as_string_with_unit(sample_rate, "Sps").replace(" ", ""))


def packets(data, packet_bytes, repeat, count):
    """What `count` successive getPacket() calls return for a file holding `data` (None = end of file)."""
    out, pos = [], 0
    for _ in range(count):
        if pos + packet_bytes <= len(data):
            out.append(data[pos:pos + packet_bytes])
            pos += packet_bytes
        elif repeat and packet_bytes <= len(data):
            out.append(data[:packet_bytes])
            pos = packet_bytes
        else:
            out.append(None)
            if repeat:
                pos = len(data)  # the re-opened stream was read to its end
    return out
