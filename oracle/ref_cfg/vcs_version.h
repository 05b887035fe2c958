#define DAV1D_VERSION "1.4.1-oracle"
