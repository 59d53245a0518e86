#pragma once  // FFmpeg headers are not in this image; the index creator includes but does not use them
