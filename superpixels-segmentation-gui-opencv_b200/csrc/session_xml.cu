// session_xml.cu -- the reference's session data file (<base>-segmentation-data.xml), host side only.
//
// MainWindow::on_button_save_session_clicked / on_button_load_session_clicked (mainwindow.cpp:673-706, 775-820) write and read
// an OpenCV FileStorage XML: GridColor, LabelsCount, per label LabelId<i> / LabelName<i> / LabelColor<i> (a Vec3b: a
// sequence of three numbers), then the CV_32SC1 matrices Labels and LabelsMask (type_id="opencv-matrix", dt "i").  This is the
// data format on the far side of the hot path: with it a label map computed by this library
// loads into the real GUI and a session saved by the GUI loads here.  No OpenCV needed; the files round-trip with
// cv::FileStorage (tests/test_session_xml.py pins both directions against cv2 4.13).
#include "common.cuh"
#include <string>
#include <vector>

namespace {

struct Parser {
    const char* p; const char* end;
    bool find(const char* s)
    {
        const size_t n = strlen(s);
        for (const char* q = p; q + n <= end; q++)
            if (memcmp(q, s, n) == 0) { p = q; return true; }
        return false;
    }
};
// the text of <name ...>TEXT</name> searched from `from`; returns false when absent
static bool element(const std::vector<char>& buf, size_t from, const std::string& name, size_t* t0, size_t* t1, size_t* after)
{
    const std::string open = "<" + name, close = "</" + name + ">";
    size_t i = from;
    for (;;) {
        const char* b = buf.data();
        Parser ps{b + i, b + buf.size()};
        if (!ps.find(open.c_str())) return false;
        size_t at = (size_t)(ps.p - b) + open.size();
        if (at < buf.size() && (b[at] == '>' || b[at] == ' ' || b[at] == '\t' || b[at] == '\n' || b[at] == '\r')) {   // not a longer name
            while (at < buf.size() && b[at] != '>') at++;
            if (at >= buf.size()) return false;
            Parser pe{b + at + 1, b + buf.size()};
            if (!pe.find(close.c_str())) return false;
            *t0 = at + 1; *t1 = (size_t)(pe.p - b); *after = *t1 + close.size();
            return true;
        }
        i = at;
    }
}
static bool parse_ints(const char* p, const char* end, long long* out, size_t want, size_t* got)
{
    size_t n = 0;
    while (p < end) {
        while (p < end && (*p == ' ' || *p == '\n' || *p == '\r' || *p == '\t')) p++;
        if (p >= end) break;
        bool neg = false;
        if (*p == '-') { neg = true; p++; } else if (*p == '+') p++;
        if (p >= end || *p < '0' || *p > '9') return false;
        long long v = 0;
        while (p < end && *p >= '0' && *p <= '9') { v = v * 10 + (*p - '0'); p++; }
        if (p < end && *p == '.') { p++; while (p < end && *p >= '0' && *p <= '9') p++; }        // "1." as some writers print
        if (n < want) out[n] = neg ? -v : v;
        n++;
    }
    *got = n;
    return true;
}
static std::string unescape(const char* p, const char* end)
{
    while (p < end && (*p == ' ' || *p == '\n' || *p == '\r' || *p == '\t')) p++;
    while (end > p && (end[-1] == ' ' || end[-1] == '\n' || end[-1] == '\r' || end[-1] == '\t')) end--;
    if (end - p >= 2 && *p == '"' && end[-1] == '"') { p++; end--; }
    std::string s;
    while (p < end) {
        if (*p == '&') {
            const struct { const char* e; char c; } tab[] = {{"&amp;", '&'}, {"&lt;", '<'}, {"&gt;", '>'}, {"&quot;", '"'}, {"&apos;", '\''}};
            bool hit = false;
            for (const auto& t : tab) {
                const size_t n = strlen(t.e);
                if ((size_t)(end - p) >= n && memcmp(p, t.e, n) == 0) { s.push_back(t.c); p += n; hit = true; break; }
            }
            if (hit) continue;
        }
        s.push_back(*p++);
    }
    return s;
}
static void escape_to(std::string& out, const char* s)
{
    bool quote = false;
    for (const char* q = s; *q; q++) if (!((*q >= '0' && *q <= '9') || (*q >= 'A' && *q <= 'Z') || (*q >= 'a' && *q <= 'z') || *q == '_')) quote = true;
    if (!*s || (*s >= '0' && *s <= '9')) quote = true;
    if (quote) out.push_back('"');
    for (const char* q = s; *q; q++) {
        switch (*q) {
            case '&': out += "&amp;"; break;
            case '<': out += "&lt;"; break;
            case '>': out += "&gt;"; break;
            case '"': out += "&quot;"; break;
            default: out.push_back(*q);
        }
    }
    if (quote) out.push_back('"');
}
// CV_32SC1 matrix in FileStorage's layout: lines of at most 72 characters, indented by 4
static void write_matrix(FILE* f, const char* name, const int32_t* m, size_t stride, int w, int h)
{
    fprintf(f, "<%s type_id=\"opencv-matrix\">\n  <rows>%d</rows>\n  <cols>%d</cols>\n  <dt>i</dt>\n  <data>\n    ", name, h, w);
    std::vector<char> buf;
    buf.reserve((size_t)w * 12 + 64);
    int col = 4;
    for (int y = 0; y < h; y++) {
        const int32_t* row = reinterpret_cast<const int32_t*>(reinterpret_cast<const char*>(m) + (size_t)y * stride);
        buf.clear();
        for (int x = 0; x < w; x++) {
            char t[16];
            int n = 0;
            long long v = row[x];
            const bool neg = v < 0;
            if (neg) v = -v;
            do { t[n++] = (char)('0' + v % 10); v /= 10; } while (v);
            if (neg) t[n++] = '-';
            const bool first = y == 0 && x == 0;
            if (!first) {
                if (col + 1 + n > 72) { buf.push_back('\n'); buf.insert(buf.end(), {' ', ' ', ' ', ' '}); col = 4; }
                else { buf.push_back(' '); col++; }
            }
            while (n) { buf.push_back(t[--n]); col++; }
        }
        fwrite(buf.data(), 1, buf.size(), f);
    }
    fprintf(f, "</data></%s>\n", name);
}

}  // namespace

extern "C" int spx_session_xml_write(const char* path, int grid_colour, const spx_label_t* list, int n_labels, const int32_t* labels,
                                     size_t labels_stride, const int32_t* labels_mask, size_t mask_stride, int width, int height) try
{
    using namespace spx;
    SPX_REQUIRE(path && labels && labels_mask && width > 0 && height > 0 && n_labels >= 0 && (n_labels == 0 || list) &&
                labels_stride >= (size_t)width * 4 && mask_stride >= (size_t)width * 4, SPX_ERR_BADARG, "session_xml_write: bad argument");
    FILE* f = fopen(path, "wb");
    SPX_REQUIRE(f, SPX_ERR_BADARG, "session_xml_write: cannot open %s for writing", path);
    fprintf(f, "<?xml version=\"1.0\"?>\n<opencv_storage>\n<GridColor>%d</GridColor>\n<LabelsCount>%d</LabelsCount>\n", grid_colour, n_labels);
    for (int i = 0; i < n_labels; i++) {
        std::string name;
        escape_to(name, list[i].name);
        fprintf(f, "<LabelId%d>%d</LabelId%d>\n<LabelName%d>%s</LabelName%d>\n<LabelColor%d>\n  %d %d %d</LabelColor%d>\n", i, list[i].id, i, i,
                name.c_str(), i, i, list[i].colour[0], list[i].colour[1], list[i].colour[2], i);
    }
    write_matrix(f, "Labels", labels, labels_stride, width, height);
    write_matrix(f, "LabelsMask", labels_mask, mask_stride, width, height);
    fprintf(f, "</opencv_storage>\n");
    const bool ok = fclose(f) == 0;
    SPX_REQUIRE(ok, SPX_ERR_INTERNAL, "session_xml_write: write to %s failed", path);
    return SPX_OK;
} SPX_API_CATCH

// Reads what the reference reads (mainwindow.cpp:775-820).  First call with labels == NULL to learn rows / cols / n_labels.
extern "C" int spx_session_xml_read(const char* path, int* grid_colour, spx_label_t* list, int list_capacity, int* n_labels, int32_t* labels,
                                    size_t labels_stride, int32_t* labels_mask, size_t mask_stride, int* rows, int* cols) try
{
    using namespace spx;
    SPX_REQUIRE(path, SPX_ERR_BADARG, "session_xml_read: path is NULL");
    FILE* f = fopen(path, "rb");
    SPX_REQUIRE(f, SPX_ERR_BADARG, "session_xml_read: cannot open %s", path);
    std::vector<char> buf;
    fseek(f, 0, SEEK_END);
    const long sz = ftell(f);
    fseek(f, 0, SEEK_SET);
    if (sz > 0) { buf.resize((size_t)sz); if (fread(buf.data(), 1, (size_t)sz, f) != (size_t)sz) buf.clear(); }
    fclose(f);
    SPX_REQUIRE(!buf.empty(), SPX_ERR_BADARG, "session_xml_read: %s is empty or unreadable", path);
    size_t t0, t1, after;
    SPX_REQUIRE(element(buf, 0, "opencv_storage", &t0, &t1, &after), SPX_ERR_BADARG, "session_xml_read: %s is not an OpenCV FileStorage XML", path);
    auto scalar = [&](const std::string& name, long long* v) -> bool {
        size_t a, b, c, got = 0;
        return element(buf, 0, name, &a, &b, &c) && parse_ints(buf.data() + a, buf.data() + b, v, 1, &got) && got == 1;
    };
    long long v = 0;
    if (grid_colour) *grid_colour = scalar("GridColor", &v) ? (int)v : 0;
    int nl = scalar("LabelsCount", &v) ? (int)v : 0;
    if (nl < 0) nl = 0;
    if (n_labels) *n_labels = nl;
    for (int i = 0; i < nl && i < list_capacity && list; i++) {
        spx_label_t& L = list[i];
        memset(&L, 0, sizeof(L));
        if (scalar("LabelId" + std::to_string(i), &v)) L.id = (int)v;
        size_t a, b, c;
        if (element(buf, 0, "LabelName" + std::to_string(i), &a, &b, &c)) {
            const std::string s = unescape(buf.data() + a, buf.data() + b);
            strncpy(L.name, s.c_str(), sizeof(L.name) - 1);
        }
        if (element(buf, 0, "LabelColor" + std::to_string(i), &a, &b, &c)) {
            size_t da = a, db = b, dc;
            std::vector<char> sub(buf.begin() + a, buf.begin() + b);
            size_t s0, s1;
            if (element(sub, 0, "data", &s0, &s1, &dc)) { da = a + s0; db = a + s1; }        // written as a 3x1 matrix (cv2 from Python)
            long long col[3] = {0, 0, 0};
            size_t got = 0;
            parse_ints(buf.data() + da, buf.data() + db, col, 3, &got);
            for (int k = 0; k < 3; k++) L.colour[k] = (uint8_t)col[k];
        }
    }
    int R = 0, Cc = 0;
    const char* names[2] = {"Labels", "LabelsMask"};
    int32_t* dsts[2] = {labels, labels_mask};
    const size_t strides[2] = {labels_stride, mask_stride};
    for (int m = 0; m < 2; m++) {
        size_t a, b, c;
        SPX_REQUIRE(element(buf, 0, names[m], &a, &b, &c), SPX_ERR_BADARG, "session_xml_read: no <%s> in %s", names[m], path);
        std::vector<char> sub(buf.begin() + a, buf.begin() + b);
        size_t s0, s1, sc, got = 0;
        long long r = 0, cc = 0;
        SPX_REQUIRE(element(sub, 0, "rows", &s0, &s1, &sc) && parse_ints(sub.data() + s0, sub.data() + s1, &r, 1, &got) && got == 1 &&
                    element(sub, 0, "cols", &s0, &s1, &sc) && parse_ints(sub.data() + s0, sub.data() + s1, &cc, 1, &got) && got == 1 && r > 0 && cc > 0,
                    SPX_ERR_BADARG, "session_xml_read: <%s> has no rows / cols", names[m]);
        SPX_REQUIRE(element(sub, 0, "dt", &s0, &s1, &sc) && unescape(sub.data() + s0, sub.data() + s1) == "i", SPX_ERR_BADARG,
                    "session_xml_read: <%s> is not a CV_32SC1 matrix", names[m]);
        if (m == 0) { R = (int)r; Cc = (int)cc; }
        SPX_REQUIRE((int)r == R && (int)cc == Cc, SPX_ERR_BADARG, "session_xml_read: Labels and LabelsMask differ in size");
        if (!dsts[m]) continue;
        SPX_REQUIRE(strides[m] >= (size_t)Cc * 4, SPX_ERR_BADARG, "session_xml_read: destination stride too small");
        SPX_REQUIRE(element(sub, 0, "data", &s0, &s1, &sc), SPX_ERR_BADARG, "session_xml_read: <%s> has no data", names[m]);
        std::vector<long long> tmp((size_t)R * Cc);
        SPX_REQUIRE(parse_ints(sub.data() + s0, sub.data() + s1, tmp.data(), tmp.size(), &got) && got == tmp.size(), SPX_ERR_BADARG,
                    "session_xml_read: <%s> holds %zu values, %zu expected", names[m], got, tmp.size());
        for (int y = 0; y < R; y++) {
            int32_t* row = reinterpret_cast<int32_t*>(reinterpret_cast<char*>(dsts[m]) + (size_t)y * strides[m]);
            for (int x = 0; x < Cc; x++) row[x] = (int32_t)tmp[(size_t)y * Cc + x];
        }
    }
    if (rows) *rows = R;
    if (cols) *cols = Cc;
    return SPX_OK;
} SPX_API_CATCH
