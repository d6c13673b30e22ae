"""archive_b200 -- B200 (sm_100a) implementation of the Inflate/Deflate and BZip2 block-codec hot path
of the Dart `archive` package, behind the package's own class names.  See DESIGN.md."""
from ._ffi import B200ZError, DartRangeError, LIB_PATH  # noqa: F401
from .codecs import BZip2Decoder, BZip2Encoder, Deflate, GZipDecoder, GZipEncoder, GZipEncoderWeb, ZLibEncoder, ZLibEncoderWeb, GZipDecoderWeb, Inflate, ZLibDecoder, ZLibDecoderWeb, inflate_buffer  # noqa: F401
from .streams import BIG_ENDIAN, LITTLE_ENDIAN, InputFileStream, InputMemoryStream, OutputFileStream, OutputMemoryStream  # noqa: F401
from .zip import Archive, ArchiveFile, ZipDecoder, ZipEncoder  # noqa: F401,E402
from .io import ZipFileEncoder, extract_archive_to_disk, extract_file_to_disk, get_input_extension  # noqa: F401,E402
