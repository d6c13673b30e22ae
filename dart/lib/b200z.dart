/// B200 back end for the codec hot path of package:archive (see INTEGRATION.md).
library archive_b200z;

export 'src/b200z_ffi.dart' show B200Z, B200ZException;
export 'src/codecs_b200.dart';
