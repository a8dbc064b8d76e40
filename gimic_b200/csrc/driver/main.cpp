// gimic-b200: command-line program over the native driver (the counterpart of `gimic [-y] gimic.inp`, src/gimic.in:25-159).
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../../include/gimic_b200_driver.h"

static void usage(FILE *f) {
    std::fputs("usage: gimic-b200 [-y|--dryrun] [-t TITLE] [-d LEVEL] [-o NAME] [-b fgimic] [--workdir DIR] [--vtk ascii|appended] [--device N | --devices all|0,1,..] [gimic.inp ...]\n"
               "  --devices: one process, one context + host thread per listed GPU (point slabs / plane rows split, nothing exchanged)\n"
               "  --cache-xdens: only convert the text XDENS named in the input(s) to the binary cache <xdens>.bin (host only)\n"
               "  one input: files are written to its directory (or --workdir), the report to stdout\n"
               "  several inputs (a current-profile scan): one device context, integrals batched into one tensor pass,\n"
               "  each report written to <input stem>.out\n", f);
}

int main(int argc, char **argv) {
    std::vector<const char *> files;
    const char *workdir = nullptr, *title = nullptr;
    int flags = 0, device = -1;
    bool multi = false, cache = false;
    std::vector<int> devices;
    for (int i = 1; i < argc; ++i) {
        const std::string a = argv[i];
        auto value = [&](const char *opt) -> const char * {
            if (i + 1 >= argc) { std::fprintf(stderr, "gimic-b200: %s needs a value\n", opt); std::exit(2); }
            return argv[++i];
        };
        if (a == "-h" || a == "--help") { usage(stdout); return 0; }
        else if (a == "-y" || a == "--dryrun") flags |= GIMIC_B200_RUN_DRYRUN;
        else if (a == "--workdir") workdir = value("--workdir");
        else if (a == "--cache-xdens") cache = true;
        // switches of the reference front end (src/gimic.in:36-57) that do not touch the hot path: accepted so that existing
        // command lines keep working (title / debug level / output base name only label the reference's own log)
        else if (a == "-t" || a == "--title") title = value("--title");
        else if (a == "-d" || a == "--debug" || a == "-o" || a == "--output") (void)value(a.c_str());
        else if (a == "-b" || a == "--backend") {
            const std::string v = value("--backend");
            if (v != "fgimic" && v != "gimic") { std::fprintf(stderr, "gimic-b200: backend '%s' is not provided (this is the fgimic path)\n", v.c_str()); return 2; }
        }
        else if (a == "--device") device = std::atoi(value("--device"));
        else if (a == "--devices") {
            const std::string v = value("--devices");
            multi = true;
            if (v != "all") {
                size_t pos = 0;
                while (pos <= v.size()) {
                    const size_t comma = v.find(',', pos);
                    const std::string tok = v.substr(pos, comma == std::string::npos ? std::string::npos : comma - pos);
                    if (tok.empty() || tok.find_first_not_of("0123456789") != std::string::npos) { std::fprintf(stderr, "gimic-b200: bad device list '%s'\n", v.c_str()); return 2; }
                    devices.push_back(std::atoi(tok.c_str()));
                    if (comma == std::string::npos) break;
                    pos = comma + 1;
                }
            }
        }
        else if (a == "--vtk") {
            const std::string v = value("--vtk");
            if (v == "appended") flags |= GIMIC_B200_RUN_VTK_APPENDED;
            else if (v != "ascii") { std::fprintf(stderr, "gimic-b200: --vtk must be ascii or appended\n"); return 2; }
        } else if (a == "--version") { std::printf("%s\n", gimic_b200_version()); return 0; }
        else if (!a.empty() && a[0] == '-') { std::fprintf(stderr, "gimic-b200: unknown option %s\n", a.c_str()); usage(stderr); return 2; }
        else files.push_back(argv[i]);
    }
    if (files.empty()) files.push_back("gimic.inp");
    if (cache) {                                     // host only: XDENS text -> <xdens>.bin for every input given
        for (const char *f : files) {
            char path[4096];
            if (gimic_b200_cache_xdens(f, workdir, path, (int)sizeof path) != 0) { std::fprintf(stderr, "gimic-b200: %s\n", gimic_b200_driver_last_error()); return 1; }
            std::printf(" binary density cache written: %s (set xdens to it)\n", path);
        }
        return 0;
    }
    int rc;
    if (files.size() > 1) rc = gimic_b200_run_scan((int)files.size(), files.data(), !devices.empty() ? devices[0] : device, flags);
    else {
        gimic_b200_run_opts o;
        std::memset(&o, 0, sizeof o);
        o.flags = flags; o.device = device; o.workdir = workdir; o.title = title;
        if (multi) { o.ndevices = devices.empty() ? -1 : (int)devices.size(); o.devices = devices.data(); }
        rc = gimic_b200_run(files[0], &o);
    }
    if (rc != 0) {
        std::fprintf(stderr, "gimic-b200: %s\n", gimic_b200_driver_last_error());
        return 1;
    }
    return 0;
}
