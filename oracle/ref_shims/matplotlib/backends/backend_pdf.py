class PdfPages:
    pass
